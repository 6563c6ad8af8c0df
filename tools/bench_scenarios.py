#!/usr/bin/env python
"""BASELINE config 4: many independent 8-agent InvPendulumBicycle scenarios in one batch
(block-diagonal pair interaction), sharded over the ranks by scenario index with NO communication.

    python tools/bench_scenarios.py [--scenarios 65536] [--per 8] [--steps 50] [--model invpendulum]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29521 tools/bench_scenarios.py

Prints one JSON line on rank 0 (secondary measurement; bench.py carries the same numbers for one GPU under
its ``extra.config4`` key).
"""
import argparse
import json
import math
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cyclistsocialforce_b200 import parameters as P  # noqa: E402
from cyclistsocialforce_b200.engine import AgentGroup, Engine, N_STATES  # noqa: E402

#: SURVEY 8d: InvPendulum agent-step, fp32 SoA: state s(6) + x(5) + flags ~ 2 x 48 B + forces 16 B
BYTES_PER_AGENT_STEP_C4 = 112.0


def run_scenarios(n_scen, per, steps, warmup, model, dev, rank=0, graph=True, split_steps=8, hbm_peak_gbs=None):
    """Step ``n_scen`` independent ``per``-agent scenarios of ``model`` on ``dev``; returns the measurements of
    this rank: ms per step (CUDA events around ``steps`` graph replays) and, from a second pass launched
    kernel by kernel, the durations of the pair kernel and of the per-agent kernel."""
    n = n_scen * per
    rng = np.random.default_rng(1000 + rank)
    L = 4.0 * math.sqrt(per)                           # L = 4 sqrt(8) per scenario
    s0 = np.zeros((n, N_STATES[model]))
    s0[:, 0], s0[:, 1] = rng.uniform(0, L, n), rng.uniform(0, L, n)
    s0[:, 2], s0[:, 3] = rng.uniform(-np.pi, np.pi, n), 5.0
    ang = s0[:, 2] + rng.uniform(-0.5, 0.5, n)
    d = 60.0 * np.arange(0, 6)
    q = np.zeros((n, 6, 3))                            # entry 0 = start position (reference convention)
    q[:, :, 0] = s0[:, 0:1] + d[None, :] * np.cos(ang)[:, None]
    q[:, :, 1] = s0[:, 1:2] + d[None, :] * np.sin(ang)[:, None]
    params = dict(twod=P.InvPendulumBicycleParameters, invpendulum=P.InvPendulumBicycleParameters,
                  planarpoint=P.PlanarPointBicycleParameters, balancingrider=P.BalancingRiderBicycleParameters)[model]()
    g = AgentGroup(model, s0, params, destqueues=list(q), dtype=torch.float32, device=dev)
    eng = Engine([g], dtype=torch.float32, device=dev, scenario_size=per, extent=2000.0, graph=graph)
    for _ in range(warmup):
        eng.step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.gpu_launches
    e0.record()
    for _ in range(steps):
        eng.step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    launches = eng.gpu_launches - l0
    # kernel split: the same step launched kernel by kernel with events between the launches
    eng.use_graph = False
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(split_steps)]
    for k in range(split_steps):
        ev[k][0].record()
        have_rep = eng._pair_and_road()
        ev[k][1].record()
        eng._agent_step(have_rep)
        ev[k][2].record()
    torch.cuda.synchronize(dev)
    eng.check_status()
    pair_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in ev)
    agent_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in ev)
    out = {"n_agents": n, "ms_total": ms, "ms_per_step": ms / steps, "gpu_launches": launches,
           "pair_kernel_ms": pair_ms, "agent_kernel_ms": agent_ms}
    if hbm_peak_gbs:
        ach = n * BYTES_PER_AGENT_STEP_C4 / (agent_ms * 1e-3) / 1e9
        out["roofline"] = {"bound": "hbm", "kernel": f"agent_kernel<float,{model.upper()},STEP>", "achieved": ach,
                           "peak": hbm_peak_gbs, "unit": "GB/s", "frac": ach / hbm_peak_gbs,
                           "bytes_per_agent_step": BYTES_PER_AGENT_STEP_C4, "kernel_ms": agent_ms,
                           "share_of_step": agent_ms / (pair_ms + agent_ms)}
    return out


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
        dist.barrier()
    n_scen = a.scenarios // world                      # this rank's scenarios (weak: none shared)
    peak = None
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                 "MEASURED_PEAKS.json"))).get("hbm_gbs"))
    except Exception:
        peak = 6650.0
    r = run_scenarios(n_scen, a.per, a.steps, a.warmup, a.model, dev, rank=rank, graph=not a.no_graph, hbm_peak_gbs=peak)
    ms = r["ms_total"]
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        total_agents = r["n_agents"] * world
        print(json.dumps({
            "metric": f"agent-steps/sec, {n_scen * world} independent {a.per}-agent {a.model} scenarios",
            "value": total_agents * a.steps / (ms * 1e-3), "unit": "agent-steps/s", "n_gpus": world,
            "steps": a.steps, "ms_per_step": ms / a.steps, "scenario_steps_per_s": n_scen * world * a.steps / (ms * 1e-3),
            "dtype": "f32 (dynamic state f64)", "scaling": "weak" if world > 1 else None,
            "pair_kernel_ms": r["pair_kernel_ms"], "agent_kernel_ms": r["agent_kernel_ms"], "roofline": r.get("roofline"),
            "config": {"workload": "BASELINE config 4", "scenarios": n_scen * world, "agents_per_scenario": a.per,
                       "communication": "none"}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
