"""Multi-GPU parity check (run under torchrun on a box with >= 2 GPUs):
the agent-range sharded crowd (payload all-gather over NCCL every step) must reproduce the
single-GPU crowd.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/check_sharded.py [--peer]

--peer: payload exchange over NVLink peer memory (csf_peer_*) inside the CUDA-graph step instead of the
NCCL all-gather.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cyclistsocialforce_b200 import parameters as P  # noqa: E402
from cyclistsocialforce_b200.distributed import PayloadExchange, PeerExchange, gather_rows_host, shard_bounds  # noqa: E402
from cyclistsocialforce_b200.engine import AgentGroup, Engine  # noqa: E402
from cyclistsocialforce_b200.synthetic import queues_with_start, synthetic_crowd  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    dev = torch.device("cuda", lr)
    ok = True
    peer = "--peer" in sys.argv          # NVLink peer-memory exchange (+ CUDA-graph step) instead of NCCL
    # f32: a tight check after a few steps (summation order differs between the decompositions) and a
    # loose one after 40 (road users crossing a field-of-view boundary one step apart: the mask is
    # discontinuous, differences of 1e-7 grow to 1e-3 in a 8,192-agent crowd)
    for n, dtype, steps in ((4099, torch.float64, 12), (8192, torch.float32, 4), (8192, torch.float32, 40)):
        s0, q = synthetic_crowd(n, seed=17, spacing=3.0)
        queues = queues_with_start(s0, q)
        extent = 2.0 * float(max(np.abs(s0[:, :2]).max(), np.abs(q[..., :2]).max())) + 1000.0
        lo, hi = shard_bounds(n, world)[rank]
        g = AgentGroup("twod", s0[lo:hi], P.InvPendulumBicycleParameters(), destqueues=list(queues[lo:hi]),
                       dtype=dtype, device=dev)
        ex = PeerExchange(n, rank, world, dtype, dev) if peer else PayloadExchange(n, rank, world)
        eng = Engine([g], dtype=dtype, device=dev, extent=extent, n_global=n, global_offset=lo, exchange=ex,
                     graph=peer, resort_every=16)
        ex(eng.payload)
        for _ in range(steps):
            eng.step()
        eng.check_status()
        got = gather_rows_host(g.states_numpy(), n, rank, world)
        if rank == 0:
            g1 = AgentGroup("twod", s0, P.InvPendulumBicycleParameters(), destqueues=list(queues), dtype=dtype,
                            device=dev)
            e1 = Engine([g1], dtype=dtype, device=dev, extent=extent, resort_every=16)
            for _ in range(steps):
                e1.step()
            ref = g1.states_numpy()
            err = float(np.abs(got - ref).max())
            tol = 1e-10 if dtype == torch.float64 else (5e-4 if steps <= 4 else 1e-2)
            print(f"sharded x{world} ({'peer memory + graph' if peer else 'NCCL all-gather'}) vs single GPU: n={n} {dtype} {steps} steps max|diff|={err:.3e} "
                  f"(tol {tol:g}) exchanges={ex.calls}", flush=True)
            ok = ok and err < tol
        if peer:
            del eng
            ex.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
