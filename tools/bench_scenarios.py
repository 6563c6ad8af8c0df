#!/usr/bin/env python
"""BASELINE config 4: many independent 8-agent InvPendulumBicycle scenarios in one batch
(block-diagonal pair interaction), sharded over the ranks by scenario index with NO communication.

    python tools/bench_scenarios.py [--scenarios 65536] [--per 8] [--steps 50] [--model invpendulum]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29521 tools/bench_scenarios.py

Prints one JSON line on rank 0 (secondary measurement, not the headline bench.py line).
"""
import argparse
import json
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cyclistsocialforce_b200 import parameters as P  # noqa: E402
from cyclistsocialforce_b200.engine import AgentGroup, Engine, N_STATES  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenarios", type=int, default=65536)
    ap.add_argument("--per", type=int, default=8)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--model", default="invpendulum")
    ap.add_argument("--no-graph", action="store_true", help="kernel-by-kernel launches (for ncu)")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n_scen = a.scenarios // world                      # this rank's scenarios (weak: none shared)
    n = n_scen * a.per
    rng = np.random.default_rng(1000 + rank)
    L = 4.0 * math.sqrt(a.per)                         # L = 4 sqrt(8) per scenario
    s0 = np.zeros((n, N_STATES[a.model]))
    s0[:, 0], s0[:, 1] = rng.uniform(0, L, n), rng.uniform(0, L, n)
    s0[:, 2], s0[:, 3] = rng.uniform(-np.pi, np.pi, n), 5.0
    ang = s0[:, 2] + rng.uniform(-0.5, 0.5, n)
    d = 60.0 * np.arange(0, 6)
    q = np.zeros((n, 6, 3))                            # entry 0 = start position (reference convention)
    q[:, :, 0] = s0[:, 0:1] + d[None, :] * np.cos(ang)[:, None]
    q[:, :, 1] = s0[:, 1:2] + d[None, :] * np.sin(ang)[:, None]
    params = dict(twod=P.InvPendulumBicycleParameters, invpendulum=P.InvPendulumBicycleParameters,
                  planarpoint=P.PlanarPointBicycleParameters, balancingrider=P.BalancingRiderBicycleParameters)[a.model]()
    g = AgentGroup(a.model, s0, params, destqueues=list(q), dtype=torch.float32, device=dev)
    eng = Engine([g], dtype=torch.float32, device=dev, scenario_size=a.per, extent=2000.0, graph=not a.no_graph)
    for _ in range(a.warmup):
        eng.step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        eng.step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    eng.check_status()
    if rank == 0:
        total_agents = n * world
        print(json.dumps({
            "metric": f"agent-steps/sec, {n_scen * world} independent {a.per}-agent {a.model} scenarios",
            "value": total_agents * a.steps / (ms * 1e-3), "unit": "agent-steps/s", "n_gpus": world,
            "steps": a.steps, "ms_per_step": ms / a.steps, "scenario_steps_per_s": n_scen * world * a.steps / (ms * 1e-3),
            "dtype": "f32 (dynamic state f64)", "scaling": "weak" if world > 1 else None,
            "config": {"workload": "BASELINE config 4", "scenarios": n_scen * world, "agents_per_scenario": a.per,
                       "communication": "none"}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
