"""Host-side logic of the N>1 path with world_size-2 gloo on CPU: agent-range sharding and the
per-step payload all-gather (the one exchange step of the sharded crowd)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cyclistsocialforce_b200.distributed import PayloadExchange, gather_rows_host, shard_bounds


def test_shard_bounds():
    assert shard_bounds(10, 1) == [(0, 10)]
    assert shard_bounds(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert shard_bounds(65536, 8)[-1] == (57344, 65536)
    for n, w in ((7, 2), (1000, 8), (3, 4)):
        b = shard_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ex = PayloadExchange(n, rank, world)
        ref = torch.arange(n * 4, dtype=torch.int32).reshape(n, 4)
        payload = torch.full((n, 4), -1, dtype=torch.int32)
        payload[ex.lo:ex.hi] = ref[ex.lo:ex.hi]           # "this rank's agent kernel wrote its rows"
        ex(payload)
        ok = bool(torch.equal(payload, ref))
        # a second step with changed local rows
        payload[ex.lo:ex.hi] += 7
        ex(payload)
        ok = ok and bool(torch.equal(payload, ref + 7))
        rows = gather_rows_host(np.full((ex.hi - ex.lo, 2), float(rank)), n, rank, world)
        ok = ok and rows.shape == (n, 2) and rows[0, 0] == 0.0 and rows[-1, 0] == float(world - 1)
        ret[rank] = (ok, ex.calls)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [64, 65])          # equal shards (all_gather_into_tensor) and ragged
def test_payload_exchange_gloo_world2(n):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, ret), nprocs=world, join=True)
    assert all(ret[r][0] for r in range(world)), dict(ret)
    assert all(ret[r][1] == 2 for r in range(world))


def test_balanced_spatial_partition():
    """Spatial decomposition helpers: Hilbert order of the initial positions, ranges balanced by the
    estimated neighbour count (host logic of bench.py's multi-GPU partition)."""
    import numpy as np
    from cyclistsocialforce_b200.distributed import balanced_bounds, neighbour_work, shard_bounds
    from cyclistsocialforce_b200.synthetic import spatial_order, synthetic_crowd
    s0, _ = synthetic_crowd(4096, seed=2, spacing=4.0)
    order = spatial_order(s0[:, 0], s0[:, 1])
    assert np.array_equal(np.sort(order), np.arange(4096))
    xs, ys = s0[order, 0], s0[order, 1]
    # consecutive agents along the curve are close: a shard is a compact region
    step = np.hypot(np.diff(xs), np.diff(ys))
    assert np.median(step) < 8.0 and step.max() < 64.0
    w = neighbour_work(xs, ys, 60.0)
    # brute-force neighbour count on a sample agrees with the cell estimate within the cell slack
    for j in (0, 1000, 4095):
        exact = (np.hypot(xs - xs[j], ys - ys[j]) <= 60.0).sum()
        assert 0.6 * exact <= w[j] <= 1.9 * exact + 0.1 * w.mean()
    b = balanced_bounds(w, 8)
    assert b[0][0] == 0 and b[-1][1] == 4096 and all(b[r][1] == b[r + 1][0] for r in range(7))
    share = np.array([w[lo:hi].sum() for lo, hi in b]) / w.sum()
    assert np.abs(share - 0.125).max() < 0.01
    assert balanced_bounds(np.ones(10), 3) == shard_bounds(10, 3) or sum(hi - lo for lo, hi in balanced_bounds(np.ones(10), 3)) == 10


def test_partition_order_keeps_a_shards_target_blocks_compact():
    """Why a sharded engine visits its targets in the order of the agent numbering: a rank's
    agents are a segment of the Hilbert curve they were numbered along, so 64 consecutive ones are always
    neighbours; re-sorted along the curve of a bounding box that has moved by a few metres the same agents are
    no longer a segment, and a few blocks of 64 consecutive ones span hundreds of metres -- the pair kernel
    cannot cull anything for such a block."""
    from cyclistsocialforce_b200.synthetic import spatial_order, synthetic_crowd

    def block_radii(x, y):
        nb = len(x) // 64
        xs, ys = x[:nb * 64].reshape(nb, 64), y[:nb * 64].reshape(nb, 64)
        return 0.5 * np.hypot(xs.max(axis=1) - xs.min(axis=1), ys.max(axis=1) - ys.min(axis=1))

    n = 65536
    s0, _ = synthetic_crowd(n, seed=1)
    o = spatial_order(s0[:, 0], s0[:, 1])
    x, y = s0[o, 0], s0[o, 1]
    worst_resorted = 0.0
    for lo, hi in ((0, 32768), (8192, 16384), (20000, 28192)):       # a half, an eighth, an unaligned eighth
        xs, ys = x[lo:hi], y[lo:hi]
        r0 = block_radii(xs, ys)
        assert r0.max() < 1.5 * np.median(r0) and np.median(r0) < 30.0      # partition order: every block compact
        for shift in (3.0, 10.0):           # the crowd's bounding box has grown by `shift` metres on one side
            xa, ya = np.r_[x, x.min() - shift], np.r_[y, y.min() - 0.3 * shift]
            oa = spatial_order(xa, ya)
            rank = np.empty(len(xa), np.int64)
            rank[oa] = np.arange(len(xa))
            loc = np.argsort(rank[lo:hi], kind="stable")             # the shard's agents along the new curve
            r1 = block_radii(xs[loc], ys[loc])
            assert abs(np.median(r1) - np.median(r0)) < 2.0          # typical blocks are as compact as before ...
            worst_resorted = max(worst_resorted, float(r1.max()))
    assert worst_resorted > 200.0                                    # ... but some span the whole region
