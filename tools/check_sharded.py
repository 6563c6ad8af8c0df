"""Multi-GPU parity check (run under torchrun on a box with >= 2 GPUs): the crowd sharded by agent range
over the ranks -- the partition bench.py uses (Hilbert order, ranges balanced by estimated neighbour count)
-- must reproduce the single-GPU crowd.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/check_sharded.py [--peer] [--big]

--peer: payload exchange over NVLink peer memory (csf_peer.cuh, folded into the step's kernels, CUDA-graph
        step) instead of the NCCL all-gather.
--big : adds the benchmark crowd itself (N = 65,536, f32).
Every case appends one JSON line to gpurun_out/sharded_check.jsonl (kept under profiles/).
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cyclistsocialforce_b200 import parameters as P  # noqa: E402
from cyclistsocialforce_b200.distributed import (PayloadExchange, PeerExchange, balanced_bounds, gather_rows_host,  # noqa: E402
                                                 neighbour_work)
from cyclistsocialforce_b200.engine import AgentGroup, Engine  # noqa: E402
from cyclistsocialforce_b200.synthetic import queues_with_start, spatial_order, synthetic_crowd  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    dev = torch.device("cuda", lr)
    ok = True
    peer = "--peer" in sys.argv          # NVLink peer-memory exchange (+ CUDA-graph step) instead of NCCL
    # f64: the sharded crowd IS the single-GPU crowd (1e-10).  f32: the order of summation depends on the
    # decomposition (a rank's pair kernel splits its items differently), so states differ by rounding after
    # one step (max-norm bound 1e-4 after 4 / 8 steps); over 40 steps a road user that crosses another's
    # field-of-view boundary one step apart in the two runs moves its neighbourhood by ~1e-3 m -- the mask is
    # discontinuous -- so the long run is held to robust statistics: median and the share of road users
    # further than 1e-4 from the single-GPU crowd.
    cases = [(4099, torch.float64, 12, "max", 1e-10), (8192, torch.float32, 4, "max", 1e-4),
             (8192, torch.float32, 40, "robust", None)]
    if "--big" in sys.argv:
        cases.append((65536, torch.float32, 8, "sparse", 1e-4))
    for n, dtype, steps, kind, tol in cases:
        s0, q = synthetic_crowd(n, seed=17, spacing=3.0 if n < 65536 else 4.0)
        order = spatial_order(s0[:, 0], s0[:, 1])
        s0, q = s0[order], q[order]
        queues = queues_with_start(s0, q)
        origin, extent = P.payload_frame([s0[:, :2], q[..., :2]])
        bounds = balanced_bounds(neighbour_work(s0[:, 0], s0[:, 1], 160.0), world)
        lo, hi = bounds[rank]
        g = AgentGroup("twod", s0[lo:hi], P.InvPendulumBicycleParameters(), destqueues=list(queues[lo:hi]),
                       dtype=dtype, device=dev)
        ex = (PeerExchange(n, rank, world, dtype, dev, bounds=bounds) if peer
              else PayloadExchange(n, rank, world, bounds=bounds))
        eng = Engine([g], dtype=dtype, device=dev, extent=extent, origin=origin, n_global=n, global_offset=lo,
                     exchange=ex, graph=peer, resort_every=16)
        ex(eng.payload)
        for _ in range(steps):
            eng.step()
        eng.check_status()
        got = gather_rows_host(g.states_numpy(), n, rank, world)
        if rank == 0:
            g1 = AgentGroup("twod", s0, P.InvPendulumBicycleParameters(), destqueues=list(queues), dtype=dtype,
                            device=dev)
            e1 = Engine([g1], dtype=dtype, device=dev, extent=extent, origin=origin, resort_every=16)
            for _ in range(steps):
                e1.step()
            ref = g1.states_numpy()
            d = np.abs(got - ref).max(axis=1)
            rec = dict(test="sharded_vs_single_gpu", world=world, exchange="peer memory, fused step, CUDA graph" if peer
                       else "NCCL all-gather", n=n, dtype=str(dtype), steps=steps, partition="balanced_bounds",
                       shard_sizes=[b - a for a, b in bounds], max_abs_diff=float(d.max()),
                       median_abs_diff=float(np.median(d)), share_over_1e4=float((d > 1e-4).mean()))
            if kind == "max":
                good = d.max() < tol
                rec["tolerance"] = tol
            elif kind == "sparse":
                # the benchmark crowd: among 65,536 road users a handful have another one within rounding
                # distance of their field-of-view boundary during the 8 steps; the two runs then disagree
                # on that one pair (the mask is discontinuous) and on nothing else
                over = int((d > tol).sum())
                good = over <= max(1, n // 8192) and np.median(d) < 1e-6 and d.max() < 0.01
                rec["tolerance"] = f"at most {max(1, n // 8192)} road users over {tol} (FOV-boundary pairs), median < 1e-6, max < 0.01"
                rec["road_users_over_tolerance"] = over
            else:
                good = np.median(d) < 5e-6 and (d > 1e-4).mean() < 0.02 and d.max() < 0.05
                rec["tolerance"] = "median < 5e-6, share(|diff| > 1e-4) < 2 %, max < 0.05"
            rec["ok"] = bool(good)
            print(json.dumps(rec), flush=True)
            out = os.path.join(ROOT, "gpurun_out")
            if os.path.isdir(out):
                with open(os.path.join(out, "sharded_check.jsonl"), "a") as f:
                    f.write(json.dumps(rec) + "\n")
            ok = ok and good
        if peer:
            del eng
            ex.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
