"""Multi-GPU driver: structures are independent, so they are sharded by structure across ranks
(one process per GPU) with NO collective on the data path; token ids are gathered to rank 0 once at the
end (4 B per token).  Replaces the reference's `jax.pmap` data parallelism
(scripts/inference_runner.py:191,299-306: batch reshaped to [Dev, B, ...], params replicated).

The partition is a greedy longest-processing-time assignment on the cost model
c(L) = L * (alpha + beta * L)  (O(L*K) encoder work + O(L^2) k-NN work), so ragged batches balance.

Callers: `InferenceRunner.tokenize` under torchrun (files sharded over ranks, rank 0 writes the token files),
`_TokenizeFn` (LPT over the devices of one process) and `bench.py --gpus N` (the BASELINE configs[4] workload).
"""
from __future__ import annotations

import heapq
from typing import Callable, List, NamedTuple, Optional, Sequence

import numpy as np


def structure_cost(length, alpha: float = 1.0, beta: float = 1.0 / 4096.0):
    """Relative cost of one structure (scalar or array of lengths)."""
    L = np.asarray(length, np.float64)
    c = L * (alpha + beta * L)
    return float(c) if c.ndim == 0 else c


def lpt_partition(lengths: Sequence[int], world: int, alpha: float = 1.0, beta: float = 1.0 / 4096.0) -> List[List[int]]:
    """Indices of the structures each rank owns; deterministic (ties broken by index, then by rank)."""
    n = len(lengths)
    cost = np.asarray(structure_cost(np.asarray(lengths, np.int64), alpha, beta), np.float64).reshape(-1)
    order = np.lexsort((np.arange(n), -cost))  # descending cost, ascending index
    heap = [(0.0, r) for r in range(world)]
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order.tolist():
        load, r = heapq.heappop(heap)
        shards[r].append(i)
        heapq.heappush(heap, (load + float(cost[i]), r))
    for s in shards:
        s.sort()
    return shards


class GatheredTokens(NamedTuple):
    """Rank 0's view of a gather: `flat[r]` holds rank r's tokens back to back in the order of `index[r]`.  On a CUDA
    group the `flat` arrays are views of a pinned staging buffer that the next gather of this process overwrites:
    consume (or copy) them first."""

    flat: List[np.ndarray]     # per rank, int32 [sum of that rank's token counts]
    index: List[np.ndarray]    # per rank, int64 [n_r] global structure indices
    counts: List[np.ndarray]   # per rank, int64 [n_r] tokens per structure
    n_total: int

    def to_list(self) -> List[np.ndarray]:
        """One uint32 array per structure, in global input order (views into the gathered buffers)."""
        out: List[Optional[np.ndarray]] = [None] * self.n_total
        for flat, idx, cnt in zip(self.flat, self.index, self.counts):
            ends = np.cumsum(cnt)
            u = flat.view(np.uint32)
            for i, a, b in zip(idx.tolist(), (ends - cnt).tolist(), ends.tolist()):
                out[i] = u[a:b]
        assert all(o is not None for o in out), "a structure's tokens never arrived"
        return out  # type: ignore[return-value]

    def total_tokens(self) -> int:
        return int(sum(int(c.sum()) for c in self.counts))


def gather_tokens_flat(local_indices: Sequence[int], local_counts: Sequence[int], local_flat, n_total: int, rank: int,
                       world: int, device=None, group=None) -> Optional[GatheredTokens]:
    """Collects every rank's tokens on rank 0 (None elsewhere).  `local_flat` is this rank's tokens back to back
    (int32; a NumPy array, or a torch tensor already on `device`, which then goes over NCCL without a host hop).
    Two small collectives (all_gather of the sizes, gather of the (index, count) tables) and ONE gather of the padded
    int32 payload: the only data-path collective of the whole tokenize job."""
    import torch
    import torch.distributed as dist

    idx = np.asarray(local_indices, np.int64)
    cnt = np.asarray(local_counts, np.int64)
    if world == 1:
        flat = local_flat.cpu().numpy() if isinstance(local_flat, torch.Tensor) else np.asarray(local_flat)
        return GatheredTokens([flat.astype(np.int32, copy=False)], [idx], [cnt], n_total)
    dev = device if device is not None else torch.device("cpu")
    on_gpu = dev.type == "cuda"
    if isinstance(local_flat, torch.Tensor):
        payload_local = local_flat.to(device=dev, dtype=torch.int32)
    else:
        host = torch.from_numpy(np.ascontiguousarray(local_flat, np.int32))
        payload_local = host.to(dev, non_blocking=False)
    sizes_t = torch.tensor([idx.size, int(payload_local.numel())], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes_t) for _ in range(world)]
    dist.all_gather(all_sizes, sizes_t, group=group)
    sizes = torch.stack(all_sizes).cpu().numpy()  # [world, 2]
    cap_n, cap_t = int(sizes[:, 0].max()), int(sizes[:, 1].max())
    # padded to the largest rank's size; the padding is never read back, so nothing is zero-filled
    table = torch.empty((2, max(cap_n, 1)), dtype=torch.int64, device=dev)
    if idx.size:
        table[:, : idx.size] = torch.from_numpy(np.stack([idx, cnt])).to(dev)
    if payload_local.numel() == max(cap_t, 1):
        payload = payload_local
    else:
        payload = torch.empty(max(cap_t, 1), dtype=torch.int32, device=dev)
        payload[: payload_local.numel()] = payload_local
    if rank == 0:
        # one flat receive buffer per kind: the per-rank views are what `gather` fills
        tables_all = torch.empty((world,) + tuple(table.shape), dtype=torch.int64, device=dev)
        bufs_all = torch.empty((world, payload.numel()), dtype=torch.int32, device=dev)
        tables, bufs = list(tables_all.unbind(0)), list(bufs_all.unbind(0))
    else:
        tables = bufs = None
    dist.gather(table, tables, dst=0, group=group)
    dist.gather(payload, bufs, dst=0, group=group)
    if rank != 0:
        return None
    if on_gpu:
        # device -> pinned host memory (pageable destinations run at a fraction of the PCIe rate), one copy per kind
        host_bufs = _pinned("tokens", (world, payload.numel()), torch.int32)
        host_tabs = _pinned("tables", (world,) + tuple(table.shape), torch.int64)
        host_bufs.copy_(bufs_all, non_blocking=True)
        host_tabs.copy_(tables_all, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        bufs_np, tabs_np = host_bufs.numpy(), host_tabs.numpy()
    else:
        bufs_np, tabs_np = bufs_all.numpy(), tables_all.numpy()
    flats, idxs, cnts = [], [], []
    for r in range(world):
        n_r, t_r = int(sizes[r, 0]), int(sizes[r, 1])
        idxs.append(tabs_np[r, 0, :n_r].copy())
        cnts.append(tabs_np[r, 1, :n_r].copy())
        flats.append(bufs_np[r, :t_r])  # a VIEW of the staging buffer: valid until this process gathers again
    return GatheredTokens(flats, idxs, cnts, n_total)


_PINNED = {}


def _pinned(kind: str, shape, dtype):
    """Grow-only pinned host staging buffers of rank 0's gather (kept between calls)."""
    import torch

    need = int(np.prod(shape))
    buf = _PINNED.get(kind)
    if buf is None or buf.numel() < need or buf.dtype != dtype:
        buf = _PINNED[kind] = torch.empty(need, dtype=dtype).pin_memory()
    return buf[:need].view(*shape)


def gather_tokens(local_indices: Sequence[int], local_tokens: Sequence[np.ndarray], n_total: int, rank: int, world: int,
                  device=None, group=None) -> Optional[List[np.ndarray]]:
    """List form of `gather_tokens_flat`: rank 0 gets one uint32 array per structure in input order."""
    counts = [int(np.asarray(t).size) for t in local_tokens]
    flat = (np.concatenate([np.asarray(t).reshape(-1) for t in local_tokens]).astype(np.int32)
            if local_tokens else np.zeros(0, np.int32))
    g = gather_tokens_flat(local_indices, counts, flat, n_total, rank, world, device=device, group=group)
    return None if g is None else g.to_list()


def tokenize_sharded(lengths: Sequence[int], tokenize_local: Callable[[List[int]], List[np.ndarray]], rank: int, world: int,
                     device=None, group=None) -> Optional[List[np.ndarray]]:
    """`tokenize_local(indices)` runs the GPU path on this rank's structures; rank 0 gets all tokens in input order."""
    shards = lpt_partition(lengths, world)
    mine = shards[rank]
    toks = tokenize_local(mine) if mine else []
    return gather_tokens(mine, toks, len(lengths), rank, world, device=device, group=group)


def dist_env():
    """(rank, world, local_rank) of this process: the torch.distributed group when one is initialised, else 0, 1, 0."""
    import os

    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), int(os.environ.get("LOCAL_RANK", "0"))
    except Exception:
        pass
    return 0, 1, 0
