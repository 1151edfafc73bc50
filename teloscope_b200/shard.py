"""Multi-GPU host logic: contigs / reads of one job sharded over the ranks of one box.

The scan has no exchange step: a contig, a contig end or a read is an independent
unit, exactly as in the reference where each is one thread-pool job (src/input.cpp:842,892,977)
whose result is appended under a mutex and sorted back into input order afterwards.  So the N-GPU
path is: partition the items by bases, every rank scans its share on its own GPU through its own
ctx, rank 0 gathers the per-item results over the host and puts them back in submission order.
No collective touches the data path; `torch.distributed` carries only the small result objects
(gloo group) and the timing barrier.

The per-rank scan is a callable so that the CPU tests can drive this module with world_size 2 over
gloo; the product passes `Scanner.scan_batch`.
"""
from __future__ import annotations

import heapq
from typing import Callable, Sequence

import numpy as np

__all__ = ["partition_items", "ShardedJob"]


def partition_items(lengths: Sequence[int], world: int) -> list[np.ndarray]:
    """Longest-processing-time-first partition of items over `world` ranks by bases.

    Returns, per rank, the ascending list of item indices it owns (ascending so that a rank's
    batch keeps the submission order of its items).  Deterministic: ties go to the lower rank,
    equal lengths are placed in index order.  Every rank computes the same answer locally."""
    if world < 1:
        raise ValueError("world must be >= 1")
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    heap = [(0, r) for r in range(world)]
    owned: list[list[int]] = [[] for _ in range(world)]
    for i in order:
        load, r = heapq.heappop(heap)
        owned[r].append(int(i))
        heapq.heappush(heap, (load + int(lengths[i]), r))
    return [np.array(sorted(o), dtype=np.int64) for o in owned]


class ShardedJob:
    """One job (a parsed assembly or a block of reads) scanned by all ranks.

    scan_fn(seqs, abs_pos, mode) -> object with .item(i) (a BatchResult, or anything shaped like
    it); every rank calls run() with the SAME item list (each rank parsed or received the same
    index; only its own share of the sequence needs to be resident)."""

    def __init__(self, scan_fn: Callable, rank: int = 0, world: int = 1, group=None):
        self.scan_fn, self.rank, self.world, self.group = scan_fn, rank, world, group

    def run(self, seqs: Sequence[bytes], abs_pos: Sequence[int] | None, mode: int, dst: int = 0):
        """-> on rank `dst`: list of per-item result dicts in submission order; elsewhere None."""
        n = len(seqs)
        parts = partition_items([len(s) for s in seqs], self.world)
        mine = parts[self.rank]
        ap = None if abs_pos is None else np.asarray([abs_pos[i] for i in mine], dtype=np.uint64)
        local = []
        if len(mine):
            res = self.scan_fn([seqs[i] for i in mine], ap, mode)
            local = [res.item(k) for k in range(len(mine))]
        if self.world == 1:
            return local
        import torch.distributed as dist
        gathered = [None] * self.world if self.rank == dst else None
        dist.gather_object(local, gathered, dst=dst, group=self.group)
        if self.rank != dst:
            return None
        out = [None] * n
        for r, items in enumerate(gathered):
            if len(items) != len(parts[r]):
                raise RuntimeError(f"rank {r} returned {len(items)} results for {len(parts[r])} items")
            for k, i in enumerate(parts[r]):
                out[int(i)] = items[k]
        return out
