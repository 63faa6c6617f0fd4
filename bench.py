#!/usr/bin/env python
"""bench.py — utterances/s of the decode hot path on synthetic [B,T,N] emissions.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path
  python bench.py --impl reference [...]                          # the reference's CPU decoder
  torchrun --nproc-per-node N bench.py --gpus N ...               # one rank per GPU, weak scaling

A step = one decode of the whole per-GPU batch (token-beam select + beam step + n-best backtrace).
`value` is timed with CUDA events on the decoder's stream with emissions already resident in HBM;
`e2e` is the same batch through the C-ABI call with HOST (pinned) emissions, the PCIe copy and the
n-best device->host copy inside the timed region. Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PEAKS_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="lexfree", choices=["lexfree", "lexicon", "lexicon_lm", "lexfree_tokenlm"],
                   help="lexfree = BASELINE configs[1]; lexicon = configs[2]; lexicon_lm = configs[3] with a synthetic 4-gram")
    p.add_argument("--lm-weight", type=float, default=2.0)
    p.add_argument("--ngrams", default="500000,500000,250000", help="2-,3-,4-gram counts of the synthetic ARPA")
    p.add_argument("--batch", type=int, default=256, help="utterances per GPU")
    p.add_argument("--frames", type=int, default=1000)
    p.add_argument("--tokens", type=int, default=10000)
    p.add_argument("--beam", type=int, default=0, help="0 = 50 (lexfree) / 100 (lexicon)")
    p.add_argument("--bst", type=int, default=0, help="beamSizeToken; 0 = N (token pruning off)")
    p.add_argument("--threshold", type=float, default=1e9)
    p.add_argument("--log-add", action="store_true", help="logAdd merging (full expansion on the device)")
    p.add_argument("--words", type=int, default=200000, help="lexicon size (lexicon workload)")
    p.add_argument("--sigma", type=float, default=1.0)
    p.add_argument("--nbest", type=int, default=0, help="0 = all final hypotheses (beam)")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-seconds", type=float, default=20.0)
    return p.parse_args()


def workload_name(a, beam, bst):
    kind = "LexiconFreeDecoder" if a.workload.startswith("lexfree") else f"LexiconDecoder {a.words}-word Trie"
    if a.log_add:
        kind += " logAdd"
    if a.workload == "lexfree_tokenlm":
        return (f"{kind}, synthetic 4-gram token ARPA ({a.ngrams} 2/3/4-grams), lmWeight={a.lm_weight}, CTC, "
                f"N={a.tokens}, T={a.frames}, beam={beam}, beamSizeToken={bst}, beamThreshold={a.threshold}, "
                f"batch={a.batch}/GPU")
    if a.workload == "lexicon_lm":
        return (f"{kind}, synthetic 4-gram ARPA ({a.ngrams} 2/3/4-grams), lmWeight={a.lm_weight}, CTC, N={a.tokens}, "
                f"T={a.frames}, beam={beam}, beamSizeToken={bst}, beamThreshold={a.threshold}, batch={a.batch}/GPU")
    thr = f"beamThreshold={a.threshold}, " if a.log_add else ""
    return (f"{kind}, ZeroLM, CTC, N={a.tokens}, T={a.frames}, beam={beam}, beamSizeToken={bst}, {thr}"
            f"batch={a.batch}/GPU")


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return PEAKS_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).
    NVML (pynvml) polls every few ms; nvidia-smi is the fallback when NVML is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _nvml_sample(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
        r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in (("hw_slowdown", n.nvmlClocksThrottleReasonHwSlowdown),
                          ("hw_thermal_slowdown", n.nvmlClocksThrottleReasonHwThermalSlowdown),
                          ("sw_thermal_slowdown", n.nvmlClocksThrottleReasonSwThermalSlowdown),
                          ("sw_power_cap", n.nvmlClocksThrottleReasonSwPowerCap)):
            if r & bit:
                self.reasons.add(name)

    def _smi_sample(self):
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True,
                             timeout=5).stdout.strip()
        if not out:
            return
        r = [x.strip() for x in out.split(",")]
        if r[0].replace(".", "").isdigit():
            self.sm.append(float(r[0]))
        if r[1].replace(".", "").isdigit():
            self.max_mhz = float(r[1])
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if r[3 + i].lower().startswith("active"):
                self.reasons.add(name)

    def run(self):
        while not self.stop_flag:
            try:
                self._nvml_sample() if self.nvml else self._smi_sample()
            except Exception:
                pass
            time.sleep(0.002 if self.nvml else 0.2)

    def summary(self):
        self.stop_flag = True
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(sm),
                "source": "nvml" if self.nvml else "nvidia-smi"}


def build_spec(a, beam, bst):
    from cases import spec_lexfree, spec_lexicon
    from text_b200 import synth

    N = a.tokens
    if a.workload == "lexfree":
        return spec_lexfree(N, beam, bst, a.threshold, sil=0, blank=N - 1, log_add=a.log_add)
    if a.workload == "lexfree_tokenlm":
        counts = [0] + [int(x) for x in a.ngrams.split(",")]
        path = os.path.join(synth.cache_dir(), f"bench4tok_{N}_{'_'.join(map(str, counts))}.arpa")
        if not os.path.exists(path):
            synth.write_arpa(path, N, order=4, counts=counts, seed=12)
        return spec_lexfree(N, beam, bst, a.threshold, sil=0, blank=N - 1, log_add=a.log_add,
                            lm_weight=a.lm_weight, lm=("arpa", path, synth.word_names(N)))
    sp = synth.lexicon(a.words, N, 2, 5, seed=7, exclude=(0, N - 1))
    if a.workload == "lexicon_lm":
        counts = [0] + [int(x) for x in a.ngrams.split(",")]
        path = os.path.join(synth.cache_dir(), f"bench4_{a.words}_{'_'.join(map(str, counts))}.arpa")
        if not os.path.exists(path):
            synth.write_arpa(path, a.words, order=4, counts=counts, seed=11)
        return spec_lexicon(N, beam, bst, sp, a.threshold, sil=0, blank=N - 1, unk=a.words, lm_weight=a.lm_weight,
                            lm=("arpa", path, synth.word_names(a.words) + ["<unk>"]), log_add=a.log_add)
    return spec_lexicon(N, beam, bst, sp, a.threshold, sil=0, blank=N - 1, unk=a.words, log_add=a.log_add)


def cpu_leg(a, spec, sample_em, seconds, kind_pref=("ref", "ora")):
    """Time the reference's decode() (or the oracle port) on all host threads over a bounded sample:
    `sample_em` [P,Ts,N]; returns (utt/s extrapolated linearly in T, description)."""
    from cases import Built
    from oracle import pyoracle as po

    kind = next((k for k in kind_pref if po.available(k)), None)
    if kind is None:
        po.build("ora")
        kind = "ora"
    O = po.Oracle(kind)
    b = Built(O, spec)
    threads = os.cpu_count() or 1
    P, Ts, N = sample_em.shape
    lex = spec["kind"] == "lexicon"
    # calibrate on one utterance, then size the sample for ~`seconds`
    t0 = time.perf_counter()
    O.bench_mt(lex, spec["opt"], b.trie, b.lm, spec["sil"], spec["blank"], spec["unk"], sample_em[:1], 1, 0)
    one = time.perf_counter() - t0
    per_thread = max(1, min(int(seconds / max(one, 1e-3)), 16))
    count = threads * per_thread
    em = sample_em[np.arange(count) % P]
    wall = O.bench_mt(lex, spec["opt"], b.trie, b.lm, spec["sil"], spec["blank"], spec["unk"], em,
                      threads, 0)
    b.close()
    frac = Ts / float(a.frames)
    ups = count / wall * frac
    desc = (f"{count} utterances x first {Ts} of {a.frames} frames on {threads} threads "
            f"({'unmodified reference compiled in place' if kind == 'ref' else 'oracle port'}); "
            + ("full length" if Ts == a.frames else "extrapolated linearly in T"))
    return ups, threads, ("reference" if kind == "ref" else "port"), desc, wall


def sample_frames(a, bst):
    # with token pruning off the reference allocates ~66 MB of LMState per frame:
    # bound the CPU sample to a short prefix there
    return a.frames if bst <= 256 else min(a.frames, 12)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from text_b200 import synth

    beam = a.beam or {"lexfree": 50, "lexicon": 100, "lexicon_lm": 200, "lexfree_tokenlm": 50}[a.workload]
    bst = a.bst or a.tokens
    spec = build_spec(a, beam, bst)
    Ts = sample_frames(a, bst)
    em = synth.emissions(4, Ts, a.tokens, seed=1234, sigma=a.sigma)
    vals = []
    for _ in range(a.warmup):
        cpu_leg(a, spec, em, 1.0)
    t_all = time.perf_counter()
    for _ in range(a.steps):
        ups, threads, kind, desc, wall = cpu_leg(a, spec, em, max(2.0, a.cpu_seconds / max(a.steps, 1)))
        vals.append(ups)
    v = float(np.mean(vals))
    out = {"impl": "reference", "metric": "utterances/sec", "value": v, "unit": "utt/s",
           "frames_per_s": v * a.frames, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": (time.perf_counter() - t_all) / max(a.steps, 1) * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": {"workload": workload_name(a, beam, bst)},
           "cpu_baseline": {"value": v, "unit": "utt/s", "cores": threads, "kind": kind, "sample": desc},
           "e2e": {"value": v, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def run_ours(a):
    import torch
    import torch.distributed as dist

    from cases import Built, assert_close_nbest, assert_same_nbest, has_ties
    from flt_backend import FltBackend
    from oracle import pyoracle as po

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    # bind this rank to the CPUs (and so, by first touch, the host memory) next to its GPU: the e2e
    # leg streams 10 GB of pinned host memory per step and crossing sockets halves the PCIe rate
    try:
        import pynvml

        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception:
        pass
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    beam = a.beam or {"lexfree": 50, "lexicon": 100, "lexicon_lm": 200, "lexfree_tokenlm": 50}[a.workload]
    bst = a.bst or a.tokens
    B, T, N = a.batch, a.frames, a.tokens
    nbest = a.nbest or beam
    spec = build_spec(a, beam, bst)
    G = FltBackend("cuda")
    G.api.device = local
    built = Built(G, spec)
    api, dec = G.api, built.dec
    api.set_nbest(dec, nbest)

    # synthetic emissions, generated where they are consumed
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    em = torch.empty((B, T, N), dtype=torch.float32, device=dev)
    for b0 in range(0, B, 32):
        z = torch.randn((min(32, B - b0), T, N), generator=gen, device=dev, dtype=torch.float32)
        em[b0:b0 + z.shape[0]] = torch.log_softmax(z * a.sigma, dim=-1)
    del z
    torch.cuda.synchronize()

    stream = torch.cuda.ExternalStream(api.stream(dec), device=dev)

    def step():
        api.decode_batch_async(dec, em.data_ptr(), B, T, N)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # one synchronous decode first: data-dependent candidate capacities (lexicon enumeration, full
    # expansion) are grown by flt_decode_batch's overflow retry and then stay for the async steps
    api.decode_batch_ptr(dec, em.data_ptr(), B, T, N)
    for _ in range(a.warmup):
        step()
    api.synchronize(dec)
    api.set_timing(dec, 1)  # CUDA events around each launch; in-kernel counters stay off while timing
    clocks = ClockSampler(local)
    clocks.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    ev0.record(stream)
    for _ in range(a.steps):
        step()
        launches += api.last_launches(dec)
    ev1.record(stream)
    barrier()
    api.synchronize(dec)
    ms_total = ev0.elapsed_time(ev1)
    # per-kernel device time of the LAST timed step (events recorded around each launch)
    last = api.last_kernel_ms(dec)
    # work / phase counters from one extra, untimed step (they cost a few hundred cycles per frame)
    api.set_timing(dec, 2)
    step()
    api.synchronize(dec)
    work = api.last_stats(dec)
    api.set_timing(dec, 0)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / a.steps
    value = world * B / (ms_step * 1e-3)
    clk = clocks.summary()

    # ---- e2e: host (pinned) emissions through the C-ABI, n-best copied back, wall clock
    e2e = None
    h2d = B * T * N * 4
    d2h = B * nbest * (T + 2) * 4 * 2 + B * nbest * 3 * 8 + B * 4
    if not a.no_e2e:
        # every rank pins its own 10 GB batch: agree first that all of them could (a rank that failed
        # alone would leave the others waiting in the collectives below)
        host, err = None, ""
        try:
            host = torch.empty((B, T, N), dtype=torch.float32, pin_memory=True)
            host.copy_(em)
        except Exception as ex:
            host, err = None, f"{type(ex).__name__}: {ex}"[:300]
        flag = torch.tensor([1.0 if host is not None else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if flag.item() < 0.5:
            e2e = {"value": None, "unit": "utt/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "error": err or "another rank could not pin its host batch"}
            host = None
        else:
            torch.cuda.synchronize()
            res = None
            for _ in range(min(a.warmup, 2)):
                api.decode_batch_ptr(dec, host.data_ptr(), B, T, N)
                res = api.nbest(dec, B, T, nbest)
            from text_b200 import shard

            def collect():
                # N > 1: the n-best blocks of every rank go to rank 0 over NCCL (device buffers, no host
                # round trip), rank 0 then reads the job's result; N = 1: plain device->host copy
                if world == 1:
                    return api.nbest(dec, B, T, nbest)
                api.synchronize(dec)
                nb = api.nbest_device(dec, B, T, nbest, beam)
                local = dict(tokens=nb["tokens"], words=nb["words"], scores=nb["scores"][:, :nbest].contiguous(),
                             counts=nb["counts"])
                return shard.gather_nbest(local, world * B, T, nbest, dev)

            barrier()
            t0 = time.perf_counter()
            for _ in range(a.steps):
                api.decode_batch_ptr(dec, host.data_ptr(), B, T, N)
                res = collect()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
            e2e = {"value": world * B * a.steps / dt, "unit": "utt/s", "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h, "ms_per_step": dt / a.steps * 1e3,
                   "note": "pinned host emissions -> flt_decode_batch (PCIe copy pipelined with the "
                           "kernels) -> " + ("flt_nbest_copy" if world == 1 else "NCCL gather of the n-best blocks to rank 0 -> host")
                           + "; bound by the host link: " f"{h2d / (dt / a.steps) / 1e9:.1f} GB/s H2D achieved per GPU"}
            del host

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- parity on a sample (outside the timed region): GPU n-best vs CPU oracle, same bits
    P = 2
    Tp = sample_frames(a, bst) if bst > 256 else T
    Tp = min(Tp, 40) if bst > 256 else Tp
    sub = em[:P, :Tp].contiguous().cpu().numpy()
    got = G.decode_batch(dec, sub, nbest)
    kind = "ref" if po.available("ref") else "ora"
    O = po.Oracle(kind)
    bo = Built(O, spec)
    exact = ties = 0
    for b in range(P):
        ro = bo.decode(sub[b], nbest)
        if has_ties(ro) or (kind == "ora" and O.tie_events(bo.dec)):
            ties += 1
            continue
        try:
            (assert_close_nbest if a.log_add else assert_same_nbest)(ro, got[b], 1e-4)
            exact += 1
        except AssertionError:
            pass
    bo.close()
    parity = {"utterances": P, "frames": Tp, "nbest": nbest, "exact_match": exact,
              "excluded_for_ties": ties, "oracle": "reference" if kind == "ref" else "port",
              "tolerance": "tokens/words bit-equal, scores abs 1e-4"}

    # ---- roofline: per frame per utterance 4N + 12*beam algorithmic bytes
    peak, peak_src = peaks()
    bytes_per_utt = T * (4 * N + 12 * beam) + nbest * (T + 2) * 8 + nbest * 24
    kernels = {}
    alg = {"token_select": B * T * (4 * N + 8 * min(beam + 3, bst)),  # row read + list written
           "beam_step": B * T * 12 * beam,                                  # back-pointer records
           "backtrace": B * nbest * (T + 2) * 8,
           # fused select + step: the whole per-frame figure of SURVEY.md §8d (row read once +
           # one back-pointer record per surviving hypothesis)
           "fused_select_step": B * T * (4 * N + 12 * beam)}
    for k, v in last.items():
        if v["launches"]:
            kernels[k] = {"ms": v["ms"], "launches": v["launches"], "algorithmic_bytes": alg[k],
                          "achieved_gbs": alg[k] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else None}
    dom = max(kernels, key=lambda k: kernels[k]["ms"]) if kernels else None
    roof = None
    if dom:
        ach = kernels[dom]["achieved_gbs"]
        # DRAM bytes per launch of this kernel from the committed ncu --set full capture of the
        # same workload (profiles/traffic.json), else null
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            ent = tj.get({"fused_select_step": "flt_k_fused", "token_select": "flt_k_topm",
                          "beam_step": "flt_k_decode"}.get(dom, dom))
            if ent and ent["workload"] == workload_name(a, beam, bst):
                traffic = ent["dram_bytes_per_launch"] / max(kernels[dom]["launches"], 1) * 1.0
        except Exception:
            traffic = None
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes"] / max(kernels[dom]["launches"], 1),
                "whole_step": {"achieved": (B * bytes_per_utt) / (ms_step * 1e-3) / 1e9,
                               "frac": (B * bytes_per_utt) / (ms_step * 1e-3) / 1e9 / peak,
                               "bytes_per_utterance": bytes_per_utt},
                "step_latency_us_per_frame": next((kernels[k]["ms"] * 1e3 / T for k in
                                                   ("fused_select_step", "beam_step") if k in kernels), None)}

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        Ts = sample_frames(a, bst)
        sample = em[:4, :Ts].contiguous().cpu().numpy()
        ups, threads, kind2, desc, _ = cpu_leg(a, spec, sample, a.cpu_seconds)
        cpu = {"value": ups, "unit": "utt/s", "cores": threads, "kind": kind2, "sample": desc}

    out = {"metric": "utterances/sec", "value": value, "unit": "utt/s", "frames_per_s": value * T,
           "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": workload_name(a, beam, bst), "emissions": f"log_softmax({a.sigma}*N(0,1)) fp32",
                      "nbest": nbest, "beamThreshold": a.threshold,
                      "l2": f"inputs ({h2d / 1e9:.2f} GB/step/GPU) exceed L2 (126 MB); no flush needed"},
           "roofline": roof, "kernels": kernels, "beam_step_work": work, "cpu_baseline": cpu, "e2e": e2e,
           "gpu_launches": launches, "clocks": clk, "parity": parity,
           "workspace_bytes": api.workspace_bytes(dec)}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
