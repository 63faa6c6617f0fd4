"""Batch sharding of the decode path over the GPUs of one node.

Utterances are independent, so the path shards with no data-path collective: rank r decodes the
contiguous block [r*ceil(B/G), ...) of the batch with its own decoder (tables replicated). The two
collectives here exist only for host-originated batches:
  scatter_emissions  rank 0 holds [B,T,N] -> every rank receives its block       (NCCL: grouped send/recv)
  gather_nbest       fixed-size n-best blocks [Bl,K,T+2]x2 int32, [Bl,K,3] f64, [Bl] int32 -> rank 0
One process per GPU, `torch.distributed` for the plumbing (backend "nccl" on GPUs; the same code is
exercised with "gloo" on CPU tensors in tests/test_shard_gloo.py with a stub decode function).
"""
import numpy as np
import torch
import torch.distributed as dist


def block(B, world, rank):
    """[lo, hi) of the batch owned by `rank`: contiguous blocks of ceil(B/world)."""
    per = (B + world - 1) // world
    lo = min(B, rank * per)
    return lo, min(B, lo + per)


def scatter_emissions(emissions, shape, device, src=0, group=None):
    """Rank `src` passes the full [B,T,N] fp32 tensor (on `device`), the others pass None; every
    rank gets its own block [Bl,T,N]. Ranks whose block is empty get a 0-row tensor."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B, T, N = shape
    lo, hi = block(B, world, rank)
    mine = torch.empty((hi - lo, T, N), dtype=torch.float32, device=device)
    if rank == src:
        reqs = []
        for r in range(world):
            rlo, rhi = block(B, world, r)
            if r == src:
                mine.copy_(emissions[rlo:rhi])
            elif rhi > rlo:
                reqs.append(dist.isend(emissions[rlo:rhi].contiguous(), dst=_global(group, r), group=group))
        for q in reqs:
            q.wait()
    elif hi > lo:
        dist.recv(mine, src=_global(group, src), group=group)
    return mine


def _global(group, r):
    """`src` / `dst` of this module are ranks INSIDE `group`; torch's point-to-point and gather calls
    take global ranks."""
    return r if group is None else dist.get_global_rank(group, r)


_pinned = {}


def _to_host(t, name):
    """device tensor -> numpy; CUDA tensors land in a page-locked buffer that is reused per (name, shape) —
    a pageable destination would cap the copy of the gathered blocks (hundreds of MB per rank) far below the link"""
    if t.device.type != "cuda":
        return t.cpu().numpy()
    key = (name, tuple(t.shape), t.dtype)
    if key not in _pinned:
        if len(_pinned) > 8:
            _pinned.clear()
        _pinned[key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h = _pinned[key]
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return h.numpy()


def gather_nbest(local, B, T, K, device, dst=0, group=None):
    """`local` = dict(tokens [Bl,K,T+2] int32, words [Bl,K,T+2] int32, scores [Bl,K,3] f64,
    counts [Bl] int32) as numpy arrays or tensors for this rank's block. Returns the same dict for
    the whole batch on rank `dst` (None elsewhere). Blocks are padded to ceil(B/world) rows so the
    collective moves fixed-size buffers. On CUDA the returned arrays are page-locked buffers owned by this module
    and overwritten by the next gather of the same shape."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = (B + world - 1) // world
    L = T + 2

    def pad(x, shape, dtype):
        t = torch.zeros(shape, dtype=dtype, device=device)
        x = torch.as_tensor(x)
        if x.numel():
            t[: x.shape[0]].copy_(x.to(device))
        return t

    parts = dict(tokens=pad(local["tokens"], (per, K, L), torch.int32),
                 words=pad(local["words"], (per, K, L), torch.int32),
                 scores=pad(local["scores"], (per, K, 3), torch.float64),
                 counts=pad(local["counts"], (per,), torch.int32))
    out = {}
    for name, t in parts.items():
        bufs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, bufs, dst=_global(group, dst), group=group)
        if rank == dst:
            rows = []
            for r in range(world):
                lo, hi = block(B, world, r)
                rows.append(bufs[r][: hi - lo])
            out[name] = _to_host(torch.cat(rows, dim=0), name)
    return out if rank == dst else None


def decode_sharded(decode_local, emissions, shape, K, device, group=None):
    """Host-originated batch on rank 0 -> n-best of the whole batch on rank 0.
    decode_local(block [Bl,T,N] tensor on `device`) -> dict as in gather_nbest (arrays for Bl rows)."""
    B, T, N = shape
    mine = scatter_emissions(emissions, shape, device, group=group)
    if mine.shape[0]:
        local = decode_local(mine)
    else:
        local = dict(tokens=np.zeros((0, K, T + 2), np.int32), words=np.zeros((0, K, T + 2), np.int32),
                     scores=np.zeros((0, K, 3), np.float64), counts=np.zeros((0,), np.int32))
    return gather_nbest(local, B, T, K, device, group=group)
