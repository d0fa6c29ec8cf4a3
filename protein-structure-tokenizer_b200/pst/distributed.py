"""Multi-GPU driver: structures are independent, so they are sharded by structure across ranks
(one process per GPU) with NO collective on the data path; token ids are gathered to rank 0 once at the
end (4 B per token).  Replaces the reference's `jax.pmap` data parallelism
(scripts/inference_runner.py:191,299-306: batch reshaped to [Dev, B, ...], params replicated).

The partition is a greedy longest-processing-time assignment on the cost model
c(L) = L * (alpha + beta * L)  (O(L*K) encoder work + O(L^2) k-NN work), so ragged batches balance.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np


def structure_cost(length: int, alpha: float = 1.0, beta: float = 1.0 / 4096.0) -> float:
    return float(length) * (alpha + beta * float(length))


def lpt_partition(lengths: Sequence[int], world: int, alpha: float = 1.0, beta: float = 1.0 / 4096.0) -> List[List[int]]:
    """Indices of the structures each rank owns; deterministic (ties broken by index)."""
    order = sorted(range(len(lengths)), key=lambda i: (-structure_cost(lengths[i], alpha, beta), i))
    load = [0.0] * world
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += structure_cost(lengths[i], alpha, beta)
    for s in shards:
        s.sort()
    return shards


def gather_tokens(local_indices: Sequence[int], local_tokens: Sequence[np.ndarray], n_total: int, rank: int, world: int,
                  device=None, group=None) -> Optional[List[np.ndarray]]:
    """Collects every rank's token arrays on rank 0 (returns None elsewhere).  One size exchange
    (all_gather of int64 counts) and one `gather` of a padded int32 payload [index, length, tokens...]."""
    import torch
    import torch.distributed as dist

    if world == 1:
        out: List[Optional[np.ndarray]] = [None] * n_total
        for i, t in zip(local_indices, local_tokens):
            out[i] = np.asarray(t, np.uint32)
        return out  # type: ignore
    dev = device if device is not None else torch.device("cpu")
    parts = []
    for i, t in zip(local_indices, local_tokens):
        t = np.asarray(t).astype(np.int64)
        parts.append(np.concatenate([[i, t.size], t]))
    flat = np.concatenate(parts).astype(np.int32) if parts else np.zeros(0, np.int32)
    n = torch.tensor([flat.size], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    cap = int(max(int(s.item()) for s in sizes))
    payload = torch.zeros(max(cap, 1), dtype=torch.int32, device=dev)
    payload[: flat.size] = torch.from_numpy(flat).to(dev)
    bufs = [torch.zeros_like(payload) for _ in range(world)] if rank == 0 else None
    dist.gather(payload, bufs, dst=0, group=group)
    if rank != 0:
        return None
    out = [None] * n_total
    for r in range(world):
        buf = bufs[r][: int(sizes[r].item())].cpu().numpy()
        pos = 0
        while pos < buf.size:
            i, m = int(buf[pos]), int(buf[pos + 1])
            out[i] = buf[pos + 2 : pos + 2 + m].astype(np.uint32)
            pos += 2 + m
    assert all(o is not None for o in out), "a structure's tokens never arrived"
    return out  # type: ignore


def tokenize_sharded(lengths: Sequence[int], tokenize_local: Callable[[List[int]], List[np.ndarray]], rank: int, world: int,
                     device=None, group=None) -> Optional[List[np.ndarray]]:
    """`tokenize_local(indices)` runs the GPU path on this rank's structures; rank 0 gets all tokens in input order."""
    shards = lpt_partition(lengths, world)
    mine = shards[rank]
    toks = tokenize_local(mine) if mine else []
    return gather_tokens(mine, toks, len(lengths), rank, world, device=device, group=group)
