"""Agent-range sharding of one crowd over several GPUs (one process per GPU).

Rank r owns agents [lo_r, hi_r) -- state, destination queues and the per-agent kernels
are local -- and evaluates its targets against *all* sources.  The one exchange step per
simulation step is an all-gather of the 16 B/agent pair payload (x, y, cos psi, sin psi)
through ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests).
Independent scenarios need no communication at all: give every rank its own ``Engine``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, world):
    """Contiguous, balanced agent ranges: [(lo, hi)] per rank."""
    base, rem = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


class PayloadExchange:
    """all-gather of the pair payload; handles unequal shard sizes."""

    def __init__(self, n_global, rank, world, group=None):
        self.bounds = shard_bounds(n_global, world)
        self.rank, self.world, self.group = rank, world, group
        self.lo, self.hi = self.bounds[rank]
        self.equal = len({hi - lo for lo, hi in self.bounds}) == 1
        self.calls = 0
        self._slab = self._mine = None

    def __call__(self, payload):
        """payload: (n_global, 4) tensor whose rows [lo, hi) were just written by this rank."""
        if self.world == 1:
            return
        self.calls += 1
        if self.equal:
            dist.all_gather_into_tensor(payload, payload[self.lo:self.hi], group=self.group)
        else:
            # ragged shards: gather max-sized slabs, then copy the valid rows out
            m = max(hi - lo for lo, hi in self.bounds)
            if self._slab is None or self._slab.device != payload.device or self._slab.dtype != payload.dtype:
                self._slab = payload.new_zeros((self.world, m, payload.shape[1]))
                self._mine = payload.new_zeros((m, payload.shape[1]))
            self._mine[:self.hi - self.lo] = payload[self.lo:self.hi]
            dist.all_gather_into_tensor(self._slab.view(-1, payload.shape[1]), self._mine, group=self.group)
            for r, (lo, hi) in enumerate(self.bounds):
                if r != self.rank:
                    payload[lo:hi] = self._slab[r, :hi - lo]


def gather_rows_host(local_rows: np.ndarray, n_global: int, rank: int, world: int, group=None):
    """Host-side helper for tests: concatenate per-rank numpy rows on every rank."""
    if world == 1:
        return local_rows
    objs = [None] * world
    dist.all_gather_object(objs, local_rows, group=group)
    return np.concatenate(objs, axis=0)
