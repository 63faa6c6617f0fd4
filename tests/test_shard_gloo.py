"""Multi-process batch sharding (text_b200/shard.py) with world size 2 and 3 over gloo on CPU: the
scatter / gather plumbing must hand every utterance to exactly one rank and put the n-best blocks
back in batch order, including ragged splits (B not divisible by the world size, B < world).
The decode itself is a deterministic stub here (no GPU in this container); the same functions run
over NCCL in bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from text_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _stub_decode(K, T):
    def f(block):
        Bl = block.shape[0]
        key = block.sum(dim=(1, 2)).double().numpy()  # identifies the utterance
        tokens = np.zeros((Bl, K, T + 2), np.int32)
        words = np.full((Bl, K, T + 2), -1, np.int32)
        scores = np.zeros((Bl, K, 3), np.float64)
        for b in range(Bl):
            tokens[b] = int(round(key[b])) % 1000
            scores[b, :, 0] = key[b] - np.arange(K)
        return dict(tokens=tokens, words=words, scores=scores, counts=np.full((Bl,), K, np.int32))
    return f


def _worker(rank, world, port, B, T, N, K, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev = torch.device("cpu")
        em = None
        if rank == 0:
            em = torch.arange(B, dtype=torch.float32).view(B, 1, 1).expand(B, T, N).contiguous() / (T * N) * 7.0
        out = shard.decode_sharded(_stub_decode(K, T), em, (B, T, N), K, dev)
        if rank == 0:
            full = _stub_decode(K, T)(em)
            ok = all(np.array_equal(out[k], full[k]) for k in ("tokens", "words", "counts")) and np.allclose(
                out["scores"], full["scores"])
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,B", [(2, 7), (2, 8), (3, 2), (2, 1)])
def test_scatter_decode_gather(world, B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, 5, 6, 3, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_block_partition_covers_batch_once():
    for B in (0, 1, 5, 8, 255, 256, 4096):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard.block(B, world, r)
                seen += list(range(lo, hi))
            assert seen == list(range(B))
