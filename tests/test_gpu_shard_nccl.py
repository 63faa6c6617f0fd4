"""The N > 1 path on hardware: text_b200.shard.decode_sharded over NCCL with one process per GPU — rank 0
scatters a host-originated batch, every rank decodes its block with the CUDA library through the C-ABI, the
n-best blocks are gathered back to rank 0 straight from the decoders' device buffers — must give, bit for bit,
what one GPU decoding the whole batch gives (utterances are independent, nothing else is
exchanged). With two GPUs on the box (`gpurun --gpus 2`) the ranks use one GPU each and NCCL; on a one-GPU box
the same test runs with two gloo ranks that both decode on cuda:0 and exchange host tensors (NCCL refuses two
ranks on one device), so the sharding logic is still checked against a real decode; bench.py checks the gathered
result of its own multi-GPU e2e leg against every rank's local one."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kind, q, one_gpu=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    from cases import Built, spec_lexfree, spec_lexicon
    from flt_backend import FltBackend
    from text_b200 import shard, synth

    ordinal = 0 if one_gpu else rank
    torch.cuda.set_device(ordinal)
    dev = torch.device("cuda", ordinal)
    if one_gpu:  # two ranks share the GPU: gloo, host tensors on the wire
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    wire = torch.device("cpu") if one_gpu else dev
    try:
        N, T, B, K = 400, 60, 11, 30  # B not divisible by the world size: ragged blocks
        if kind == "lexfree":
            spec = spec_lexfree(N, K, N, 1e9)
        else:
            spec = spec_lexicon(N, K, N, synth.lexicon(3000, N, 2, 4, seed=7, exclude=(0, N - 1)), 25.0, word_score=0.3)
        G = FltBackend("cuda")
        G.api.device = ordinal
        built = Built(G, spec)
        api, dec = G.api, built.dec

        def decode_local(block):
            Bl = block.shape[0]
            if one_gpu:  # the block arrived as a host tensor: flt_decode_batch takes host pointers too
                block = block.contiguous()
                api.decode_batch_ptr(dec, block.data_ptr(), Bl, T, N)
                return api.nbest(dec, Bl, T, K)
            api.decode_batch_ptr(dec, block.data_ptr(), Bl, T, N)
            api.synchronize(dec)
            nb = api.nbest_device(dec, Bl, T, K, K)  # device tensors aliasing the decoder's buffers
            return dict(tokens=nb["tokens"], words=nb["words"], scores=nb["scores"], counts=nb["counts"])

        em = None
        if rank == 0:
            em = torch.from_numpy(synth.emissions(B, T, N, seed=5, sigma=2.0)).to(wire)
        out = shard.decode_sharded(decode_local, em, (B, T, N), K, wire)
        if rank == 0:
            em = em.to(dev)
            api.decode_batch_ptr(dec, em.data_ptr(), B, T, N)
            full = api.nbest(dec, B, T, K)
            ok = True
            for b in range(B):
                n = int(full["counts"][b])
                ok = ok and int(out["counts"][b]) == n and n > 0
                ok = ok and np.array_equal(out["tokens"][b, :n], full["tokens"][b, :n])
                ok = ok and np.array_equal(out["words"][b, :n], full["words"][b, :n])
                ok = ok and np.array_equal(out["scores"][b, :n], full["scores"][b, :n])
            q.put(bool(ok))
        built.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["lexfree", "lexicon"])
def test_nccl_sharded_decode_equals_one_gpu(kind):
    import torch
    import torch.multiprocessing as mp

    one_gpu = torch.cuda.device_count() < 2
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, q, one_gpu)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_one_trie_and_lm_shared_by_decoders_on_two_devices():
    """include/flt_decoder.h: a Trie / LM may be shared by several decoders, also across CUDA devices — the
    flattened tables are built once per device (they used to be built once, on the first decoder's
    device). Same process, one lexicon + n-gram LM, one decoder per GPU: identical n-best; and the caller's
    current device is what it was before each call."""
    import torch

    second = 1 if torch.cuda.device_count() >= 2 else 0  # one GPU: two decoders share the tables on it
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity_cases
    from flt_backend import FltBackend

    name, spec, em = next(c for c in parity_cases.lexicon_cases() if c[0] == "arpa3_ctc")
    G = FltBackend("cuda")
    api = G.api
    lm = api.lm_arpa(spec["lm"][1], spec["lm"][2])
    trie = api.trie_create(spec["N"], spec["sil"])
    for sp, wid in zip(spec["spellings"], spec["word_ids"]):
        api.trie_insert(trie, sp, wid, float(api.lm_score_seq(lm, [wid])[0]))
    api.trie_smear(trie, spec["smear"])
    torch.cuda.set_device(0)
    res = []
    for device in (0, second, 0):
        api.device = device
        dec = G.decoder_lexicon(spec["opt"], trie, lm, spec["sil"], spec["blank"], spec["unk"])
        res.append(G.decode_batch(dec, em, spec["opt"].beamSize))
        assert torch.cuda.current_device() == 0, "an ABI call left another device current"
        api.decoder_destroy(dec)
    for other in res[1:]:
        for a, b in zip(res[0], other):
            assert a["n"] == b["n"] and a["n"] > 0
            assert np.array_equal(a["tokens"], b["tokens"]) and np.array_equal(a["words"], b["words"])
            assert np.array_equal(a["scores"], b["scores"])
    api.trie_destroy(trie)
    api.lm_destroy(lm)
