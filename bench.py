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
    p.add_argument("--no-secondary", action="store_true", help="skip the cfg 3 (lexicon) block of the default run")
    p.add_argument("--no-tertiary", action="store_true",
                   help="skip the cfg 4 (lexicon + 4-gram, beam 200, T=1500, B=512) block of the default run")
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


def config_of(a, beam, bst, nbest):
    """the workload description both arms print (identical dicts: the driver compares them)"""
    h2d = a.batch * a.frames * a.tokens * 4
    return {"workload": workload_name(a, beam, bst), "emissions": f"log_softmax({a.sigma}*N(0,1)) fp32",
            "nbest": nbest, "beamThreshold": a.threshold,
            "l2": f"inputs ({h2d / 1e9:.2f} GB/step/GPU) exceed L2 (126 MB); no flush needed"}


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
            synth.write_arpa(path + f".tmp{os.getpid()}", N, order=4, counts=counts, seed=12)
            os.replace(path + f".tmp{os.getpid()}", path)
        return spec_lexfree(N, beam, bst, a.threshold, sil=0, blank=N - 1, log_add=a.log_add,
                            lm_weight=a.lm_weight, lm=("arpa", path, synth.word_names(N)))
    sp = synth.lexicon(a.words, N, 2, 5, seed=7, exclude=(0, N - 1))
    if a.workload == "lexicon_lm":
        counts = [0] + [int(x) for x in a.ngrams.split(",")]
        path = os.path.join(synth.cache_dir(), f"bench4_{a.words}_{'_'.join(map(str, counts))}.arpa")
        if not os.path.exists(path):
            synth.write_arpa(path + f".tmp{os.getpid()}", a.words, order=4, counts=counts, seed=11)
            os.replace(path + f".tmp{os.getpid()}", path)
        return spec_lexicon(N, beam, bst, sp, a.threshold, sil=0, blank=N - 1, unk=a.words, lm_weight=a.lm_weight,
                            lm=("arpa", path, synth.word_names(a.words) + ["<unk>"]), log_add=a.log_add)
    return spec_lexicon(N, beam, bst, sp, a.threshold, sil=0, blank=N - 1, unk=a.words, log_add=a.log_add)


def cpu_leg(a, spec, sample_em, seconds, kind_pref=("ref", "ora"), min_per_thread=8):
    """Time the reference's decode() (or the oracle port) on all host threads over a bounded sample:
    `sample_em` [P,Ts,N]; one decoder object per thread, one warm-up utterance per thread, then at least
    `min_per_thread` utterances per thread inside the timed region (steady state: decodeBegin's teardown of
    the previous utterance is included). Returns utt/s extrapolated linearly in T."""
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
    per_thread = max(min_per_thread, min(int(seconds / max(one, 1e-3)), 32))
    count = threads * per_thread
    em = sample_em[np.arange(count) % P]
    wall = O.bench_mt(lex, spec["opt"], b.trie, b.lm, spec["sil"], spec["blank"], spec["unk"], em,
                      threads, 1)
    b.close()
    frac = Ts / float(a.frames)
    ups = count / wall * frac
    desc = (f"{count} utterances ({per_thread} per thread after 1 warm-up each) x first {Ts} of {a.frames} frames on "
            f"{threads} threads ({'unmodified reference compiled in place' if kind == 'ref' else 'oracle port'}); "
            + ("full length" if Ts == a.frames else "extrapolated linearly in T"))
    return ups, threads, ("reference" if kind == "ref" else "port"), desc, wall


def sample_frames(a, bst):
    # with token pruning off the reference allocates ~66 MB of LMState per frame:
    # bound the CPU sample to a short prefix there
    return a.frames if bst <= 256 else min(a.frames, 12)


def default_beam(a):
    return a.beam or {"lexfree": 50, "lexicon": 100, "lexicon_lm": 200, "lexfree_tokenlm": 50}[a.workload]


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from text_b200 import synth

    beam = default_beam(a)
    bst = a.bst or a.tokens
    spec = build_spec(a, beam, bst)
    Ts = sample_frames(a, bst)
    em = synth.emissions(8, Ts, a.tokens, seed=1234, sigma=a.sigma)
    vals, walls = [], []
    for _ in range(a.warmup):
        cpu_leg(a, spec, em, 1.0, min_per_thread=1)
    for _ in range(a.steps):
        ups, threads, kind, desc, wall = cpu_leg(a, spec, em, max(2.0, a.cpu_seconds / max(a.steps, 1)))
        vals.append(ups)
        walls.append(wall)
    v = float(np.mean(vals))
    # the reference's favourable setting next to it (token beam = beam, full length), as in the main arm
    fav = None
    if bst > beam and not a.no_cpu_baseline:
        spec_b = build_spec(a, beam, beam)
        em_b = synth.emissions(8, a.frames, a.tokens, seed=1234, sigma=a.sigma)
        ups_b, thr_b, kind_b, desc_b, _ = cpu_leg(a, spec_b, em_b, a.cpu_seconds / 2, min_per_thread=3)
        fav = {"value": ups_b, "unit": "utt/s", "cores": thr_b, "kind": kind_b, "sample": desc_b, "beamSizeToken": beam}
    out = {"impl": "reference", "metric": "utterances/sec", "value": v, "unit": "utt/s",
           "frames_per_s": v * a.frames, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": float(np.mean(walls)) * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": config_of(a, beam, bst, a.nbest or beam),
           "timing": "wall clock of the multi-threaded decode of a bounded sample (see cpu_baseline.sample)",
           "cpu_baseline": {"value": v, "unit": "utt/s", "cores": threads, "kind": kind, "sample": desc,
                            "note": "one step = the timed multi-threaded decode of that sample; beamSizeToken = N is "
                                    "the CPU-hostile setting (the reference expands beam x N candidates per frame), "
                                    "see cpu_baseline_bst_beam of the main arm for the CPU-favourable one"},
           "cpu_baseline_bst_beam": fav,
           "e2e": {"value": v, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def src_hash():
    """sha256 over the device code of the kernels (the headers the __global__ wrappers instantiate): ties a
    committed ncu capture (profiles/traffic.json) to the build. Host-side files (flt_abi.cu, table_io.h,
    runtime.h, host/) are left out so that an ABI change does not make the capture look stale."""
    import hashlib

    h = hashlib.sha256()
    d = os.path.join(ROOT, "text_b200", "csrc")
    for f in ("spmd.h", "tables.h", "topm_core.h", "topm_stream.h", "beam_core.h", "beam_lf.h", "beam_gx.h",
              "fused_core.h"):
        with open(os.path.join(d, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def oracle_parity(a, G, dec, spec, em, nbest, bst, P=8):
    """GPU n-best vs the CPU oracle on the first P utterances of this rank's batch, same bits in: token /
    word rows bit-equal, scores within 1e-4. Full length where the reference can run it (bst <= 256);
    a >= 40-frame prefix with token pruning off (the port: the reference allocates ~66 MB per frame there)."""
    from concurrent.futures import ThreadPoolExecutor

    from cases import Built, assert_close_nbest, assert_same_nbest, has_ties
    from oracle import pyoracle as po

    T = a.frames
    Tp = T if bst <= 256 else min(T, 40)
    P = min(P, em.shape[0])
    sub = em[:P, :Tp].contiguous().cpu().numpy()
    got = G.decode_batch(dec, sub, nbest)
    kind = "ref" if (po.available("ref") and bst <= 256) else "ora"
    if not po.available(kind):
        po.build("ora")
        kind = "ora"
    O = po.Oracle(kind)
    A = po.Oracle("ora") if po.available("ora") else None  # tie detector

    def one(b):
        bo = Built(O, spec)
        ro = bo.decode(sub[b], nbest)
        ties = has_ties(ro) or (kind == "ora" and O.tie_events(bo.dec) > 0)
        bo.close()
        if not ties and kind == "ref" and A is not None:
            ba = Built(A, spec)
            ba.decode(sub[b], nbest)
            ties = A.tie_events(ba.dec) > 0
            ba.close()
        return ro, ties

    exact = ties = 0
    with ThreadPoolExecutor(max_workers=min(P, os.cpu_count() or 1)) as ex:
        for b, (ro, tie) in enumerate(ex.map(one, range(P))):
            if tie:
                ties += 1
                continue
            try:
                (assert_close_nbest if a.log_add else assert_same_nbest)(ro, got[b], 1e-4)
                exact += 1
            except AssertionError:
                pass
    return {"utterances": P, "frames": Tp, "nbest": nbest, "exact_match": exact, "excluded_for_ties": ties,
            "mismatch": P - exact - ties, "oracle": "reference" if kind == "ref" else "port",
            "tolerance": "tokens/words bit-equal, scores abs 1e-4"}


def measure(a, ctx, with_cpu):
    """One workload on this rank's batch: device-timed steps, per-kernel times, e2e through the C-ABI with host
    emissions, parity against the oracle, roofline. `ctx` carries the shared emissions / ranks."""
    import torch
    import torch.distributed as dist

    from cases import Built
    from text_b200 import shard

    world, rank, local, dev, em, G = ctx["world"], ctx["rank"], ctx["local"], ctx["dev"], ctx["em"], ctx["G"]
    beam = default_beam(a)
    bst = a.bst or a.tokens
    B, T, N = a.batch, a.frames, a.tokens
    nbest = a.nbest or beam
    spec = build_spec(a, beam, bst)
    G.setup_times = {}
    t0 = time.perf_counter()
    built = Built(G, spec)
    setup = {"build_tables_s": time.perf_counter() - t0,
             "note": "build_tables_s = LM load + one Trie insert per lexicon word + smear through the C-ABI from Python; "
                     "*_file_load_s = the same objects from their table files (csrc/table_io.h); first_decode_s = "
                     "planning + flattening / upload of the tables to HBM + one decode of the batch"}
    api, dec = G.api, built.dec
    if built.trie is not None and rank == 0:
        from text_b200 import synth

        tp = os.path.join(synth.cache_dir(), f"bench_trie_{os.getpid()}.flt")
        t0 = time.perf_counter()
        api.trie_save(built.trie, tp)
        setup["trie_file_save_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        t2 = api.trie_load(tp)
        setup["trie_file_load_s"] = time.perf_counter() - t0
        setup["trie_file_bytes"] = os.path.getsize(tp)
        setup["trie_nodes"] = api.trie_num_nodes(t2)
        api.trie_destroy(t2)
        os.remove(tp)
    api.set_nbest(dec, nbest)
    stream = torch.cuda.ExternalStream(api.stream(dec), device=dev)

    def step():
        api.decode_batch_async(dec, em.data_ptr(), B, T, N)

    def barrier():
        torch.cuda.synchronize()
        if world > 1 or os.environ.get("BENCH_FORCE_PG"):
            dist.barrier()
        torch.cuda.synchronize()

    # one synchronous decode first: data-dependent candidate capacities (lexicon enumeration, full
    # expansion) are grown by flt_decode_batch's overflow retry and then stay for the async steps
    t0 = time.perf_counter()
    api.decode_batch_ptr(dec, em.data_ptr(), B, T, N)
    setup["first_decode_s"] = time.perf_counter() - t0
    setup.update(G.setup_times)
    for _ in range(a.warmup):
        step()
    api.synchronize(dec)
    api.set_timing(dec, 1)  # CUDA events around each launch; in-kernel counters stay off while timing
    clocks = ClockSampler(local)
    if not os.environ.get("BENCH_NO_CLOCKS"):
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
    clk = clocks.summary()  # the sampler stops here: it only runs across the device-timed region
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

    # ---- e2e: host (pinned) emissions through the C-ABI, n-best copied back, wall clock
    e2e = None
    h2d = B * T * N * 4
    d2h = B * nbest * (T + 2) * 4 * 2 + B * nbest * 3 * 8 + B * 4
    host = ctx.get("host")
    if not a.no_e2e:
        if host is None:
            e2e = {"value": None, "unit": "utt/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "error": ctx.get("host_error") or "host batch could not be pinned"}
        else:
            torch.cuda.synchronize()
            res = None

            def collect():
                # N > 1: the n-best blocks of every rank go to rank 0 over NCCL (device buffers, no host round
                # trip), rank 0 then reads the job's result; N = 1: plain device->host copy
                if world == 1:
                    return api.nbest(dec, B, T, nbest, pinned=True)  # page-locked result buffers, like the input
                api.synchronize(dec)
                nb = api.nbest_device(dec, B, T, nbest, beam)
                mine = dict(tokens=nb["tokens"], words=nb["words"], scores=nb["scores"][:, :nbest].contiguous(),
                            counts=nb["counts"])
                return shard.gather_nbest(mine, world * B, T, nbest, dev)

            # warm-up through the same path as the timed steps: the page-locked result buffers and the NCCL
            # channels of the gather are set up here, not inside the timed region
            for _ in range(min(a.warmup, 2)):
                api.decode_batch_ptr(dec, host.data_ptr(), B, T, N)
                res = collect()
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
            # the gathered job result is checked, not discarded: every rank's block must equal what that rank
            # reads from its own decoder (CRC of tokens / words / scores / counts, exchanged out of band)
            gathered_ok = None
            if world > 1:
                import zlib

                loc = api.nbest(dec, B, T, nbest)
                crc = zlib.crc32(loc["tokens"].tobytes() + loc["words"].tobytes() + loc["scores"].tobytes()
                                 + loc["counts"].tobytes())
                crcs = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
                dist.all_gather(crcs, torch.tensor([crc], dtype=torch.int64, device=dev))
                if rank == 0:
                    gathered_ok = True
                    for r in range(world):
                        blk = slice(r * B, (r + 1) * B)
                        c2 = zlib.crc32(np.ascontiguousarray(res["tokens"][blk]).tobytes()
                                        + np.ascontiguousarray(res["words"][blk]).tobytes()
                                        + np.ascontiguousarray(res["scores"][blk]).tobytes()
                                        + np.ascontiguousarray(res["counts"][blk]).tobytes())
                        gathered_ok = gathered_ok and (c2 == int(crcs[r].item()))
            e2e = {"value": world * B * a.steps / dt, "unit": "utt/s", "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h, "ms_per_step": dt / a.steps * 1e3,
                   "h2d_gbs_per_gpu": h2d / (dt / a.steps) / 1e9,
                   "host_link": ctx.get("host_link"),
                   "gathered_result_equals_local": gathered_ok,
                   "note": "pinned host emissions -> flt_decode_batch (PCIe copy pipelined with the "
                           "kernels) -> " + ("flt_nbest_copy" if world == 1 else "NCCL gather of the n-best blocks to rank 0 -> host")
                           + "; bound by the host link (see host_link: plain pinned cudaMemcpyAsync ceiling of this box)"}

    out = {"value": value, "ms_per_step": ms_step, "launches": launches, "clocks": clk, "e2e": e2e,
           "workload": workload_name(a, beam, bst), "beam": beam, "bst": bst, "nbest": nbest, "setup": setup}
    if rank != 0:
        built.close()
        return out

    # ---- parity on a sample (outside the timed region): GPU n-best vs CPU oracle, same bits
    parity = oracle_parity(a, G, dec, spec, em, nbest, bst)

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
    # the kernel the roofline is quoted for: the one that moves the path's HBM bytes (the select, fused or
    # not); the latency-bound step is reported through whole_step
    hbm = [k for k in ("fused_select_step", "token_select") if k in kernels]
    dom = hbm[0] if hbm else (max(kernels, key=lambda k: kernels[k]["ms"]) if kernels else None)
    roof = None
    if dom:
        ach = kernels[dom]["achieved_gbs"]
        # DRAM bytes per launch of this kernel from the committed ncu --set full capture of the same workload
        # and the same kernel sources (profiles/traffic.json carries their hash), else null
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            ent = tj.get({"fused_select_step": "flt_k_fused", "token_select": "flt_k_topm",
                          "beam_step": "flt_k_decode"}.get(dom, dom))
            if ent and ent["workload"] == workload_name(a, beam, bst):
                if ent.get("src_sha") == src_hash():
                    traffic = ent["dram_bytes_per_launch"] / max(kernels[dom]["launches"], 1) * 1.0
                    traffic_src = f"committed ncu --set full capture ({ent.get('file', 'profiles/')}), kernel sources {ent['src_sha']}"
                else:
                    traffic_src = "stale: the kernel sources changed since the committed ncu capture"
        except Exception:
            traffic = None
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes"] / max(kernels[dom]["launches"], 1),
                "whole_step": {"achieved": (B * bytes_per_utt) / (ms_step * 1e-3) / 1e9,
                               "frac": (B * bytes_per_utt) / (ms_step * 1e-3) / 1e9 / peak,
                               "bytes_per_utterance": bytes_per_utt},
                "step_latency_us_per_frame": next((kernels[k]["ms"] * 1e3 / T for k in
                                                   ("fused_select_step", "beam_step") if k in kernels), None)}

    cpu = cpu_beam = None
    if with_cpu and world == 1 and not a.no_cpu_baseline:
        Ts = sample_frames(a, bst)
        sample = em[:8, :Ts].contiguous().cpu().numpy()
        ups, threads, kind2, desc, _ = cpu_leg(a, spec, sample, a.cpu_seconds)
        cpu = {"value": ups, "unit": "utt/s", "cores": threads, "kind": kind2, "sample": desc,
               "beamSizeToken": bst}
        if bst > beam:
            # the CPU-favourable setting (token beam = beam): full length on the reference
            spec_b = build_spec(a, beam, beam)
            sample = em[:8].contiguous().cpu().numpy()
            ups, threads, kind2, desc, _ = cpu_leg(a, spec_b, sample, a.cpu_seconds / 2, min_per_thread=3)
            cpu_beam = {"value": ups, "unit": "utt/s", "cores": threads, "kind": kind2, "sample": desc,
                        "beamSizeToken": beam,
                        "note": "the reference at beamSizeToken = beamSize, its favourable setting; the device path "
                                "is timed with token pruning off (beamSizeToken = N), which costs it nothing"}
    out.update({"roofline": roof, "kernels": kernels, "beam_step_work": work, "cpu_baseline": cpu,
                "cpu_baseline_bst_beam": cpu_beam, "parity": parity, "workspace_bytes": api.workspace_bytes(dec)})
    built.close()
    return out


def host_link_ceiling(dev, host, world):
    """Plain pinned-host -> device cudaMemcpyAsync rate of this box with every rank copying at once (two
    copies in flight per rank): the roof of the e2e leg, whose input is 4 N bytes per frame over this link."""
    import torch
    import torch.distributed as dist

    n = min(host.numel(), (2 << 30) // 4)
    src = host.view(-1)[:n]
    dst = [torch.empty(n // 2, dtype=torch.float32, device=dev) for _ in range(2)]
    st = [torch.cuda.Stream(device=dev) for _ in range(2)]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        for k in range(2):
            with torch.cuda.stream(st[k]):
                dst[k].copy_(src[k * (n // 2):(k + 1) * (n // 2)], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = reps * n * 4 / dt / 1e9
    t = torch.tensor([gbs], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return {"h2d_gbs_per_gpu_all_ranks_copying": float(t.item()), "ranks": world,
            "how": "pinned cudaMemcpyAsync, 2 streams per rank, 3 x 2 GiB, min over ranks"}


def run_ours(a):
    import copy

    import torch
    import torch.distributed as dist

    from flt_backend import FltBackend

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
    if world > 1 or os.environ.get("BENCH_FORCE_PG"):  # BENCH_FORCE_PG: a 1-rank NCCL group, to study its side effects
        if world == 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B, T, N = a.batch, a.frames, a.tokens
    G = FltBackend("cuda")
    G.api.device = local
    G.table_cache = True  # n-gram tables are parsed from ARPA text once and then loaded from their table file

    # synthetic emissions, generated where they are consumed
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    em = torch.empty((B, T, N), dtype=torch.float32, device=dev)
    for b0 in range(0, B, 32):
        z = torch.randn((min(32, B - b0), T, N), generator=gen, device=dev, dtype=torch.float32)
        em[b0:b0 + z.shape[0]] = torch.log_softmax(z * a.sigma, dim=-1)
    del z
    torch.cuda.synchronize()
    ctx = dict(world=world, rank=rank, local=local, dev=dev, em=em, G=G, host=None)
    if not a.no_e2e:
        # every rank pins its own batch: agree first that all of them could (a rank that failed alone
        # would leave the others waiting in the collectives of the e2e leg)
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
            host = None
            ctx["host_error"] = err or "another rank could not pin its host batch"
        ctx["host"] = host
        if host is not None:
            ctx["host_link"] = host_link_ceiling(dev, host, world)

    if world > 1:
        # synthetic lexicon / ARPA files are cached on disk: rank 0 writes them, the others wait and read
        if rank == 0:
            build_spec(a, default_beam(a), a.bst or a.tokens)
        dist.barrier()
    main = measure(a, ctx, with_cpu=True)
    # ---- secondary: BASELINE configs[2] (LexiconDecoder, 200k-word Trie, ZeroLM, beam 100), the north star's
    # target configuration, on the same emissions — so that its numbers are driver-run too
    secondary = None
    if a.workload == "lexfree" and not a.no_secondary and not a.log_add:
        a2 = copy.copy(a)
        a2.workload, a2.beam, a2.nbest = "lexicon", 0, 0
        a2.steps, a2.warmup = max(2, min(a.steps, 3)), max(1, min(a.warmup, 3))
        sec = measure(a2, ctx, with_cpu=True)
        if rank == 0:
            secondary = {"config": {"workload": sec["workload"]}, "metric": "utterances/sec", "value": sec["value"],
                         "unit": "utt/s", "ms_per_step": sec["ms_per_step"], "steps": a2.steps, "warmup": a2.warmup,
                         "kernels": sec["kernels"], "roofline": sec["roofline"], "parity": sec["parity"],
                         "e2e": sec["e2e"], "cpu_baseline": sec["cpu_baseline"],
                         "cpu_baseline_bst_beam": sec["cpu_baseline_bst_beam"], "beam_step_work": sec["beam_step_work"],
                         "gpu_launches": sec["launches"], "clocks": sec["clocks"], "setup": sec["setup"]}
    # ---- tertiary: BASELINE configs[3] (LexiconDecoder, 200k-word Trie + 4-gram LM, beam 200, beamThreshold 25,
    # T=1500, B=512 per GPU) with a synthetic ARPA of the SURVEY's size (2M/2M/1M 2/3/4-grams): device-timed
    tertiary = None
    if a.workload == "lexfree" and not a.no_tertiary and not a.no_secondary and not a.log_add:
        a3 = copy.copy(a)
        a3.workload, a3.beam, a3.nbest, a3.threshold = "lexicon_lm", 200, 0, 25.0
        a3.batch, a3.frames, a3.ngrams = 512, 1500, "2000000,2000000,1000000"
        a3.steps, a3.warmup, a3.no_e2e = 2, 1, True
        del em, ctx["em"]
        torch.cuda.empty_cache()
        if world > 1:
            if rank == 0:
                build_spec(a3, 200, a3.tokens)
            dist.barrier()
        em3 = torch.empty((a3.batch, a3.frames, N), dtype=torch.float32, device=dev)
        for b0 in range(0, a3.batch, 16):
            z = torch.randn((min(16, a3.batch - b0), a3.frames, N), generator=gen, device=dev, dtype=torch.float32)
            em3[b0:b0 + z.shape[0]] = torch.log_softmax(z * a.sigma, dim=-1)
        del z
        ctx3 = dict(ctx, em=em3, host=None)
        ter = measure(a3, ctx3, with_cpu=False)
        if rank == 0:
            tertiary = {"config": {"workload": ter["workload"]}, "metric": "utterances/sec", "value": ter["value"],
                        "unit": "utt/s", "ms_per_step": ter["ms_per_step"], "steps": a3.steps, "warmup": a3.warmup,
                        "kernels": ter["kernels"], "roofline": ter["roofline"], "parity": ter["parity"],
                        "beam_step_work": ter["beam_step_work"], "gpu_launches": ter["launches"],
                        "clocks": ter["clocks"], "setup": ter["setup"], "workspace_bytes": ter["workspace_bytes"]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    h2d = B * T * N * 4
    out = {"metric": "utterances/sec", "value": main["value"], "unit": "utt/s", "frames_per_s": main["value"] * T,
           "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": main["ms_per_step"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": config_of(a, main["beam"], main["bst"], main["nbest"]),
           "timing": "CUDA events on the decoder's stream around exactly K steps, max over ranks",
           "roofline": main["roofline"], "kernels": main["kernels"], "beam_step_work": main["beam_step_work"],
           "cpu_baseline": main["cpu_baseline"], "cpu_baseline_bst_beam": main["cpu_baseline_bst_beam"],
           "e2e": main["e2e"], "gpu_launches": main["launches"], "clocks": main["clocks"], "parity": main["parity"],
           "workspace_bytes": main["workspace_bytes"], "setup": main["setup"], "secondary": secondary,
           "tertiary": tertiary}
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
