#!/usr/bin/env python
"""Step time of one open-plane crowd per model class (secondary measurement for DESIGN.md):
K1 + K2/K3 of TwoDBicycle, InvPendulumBicycle, BalancingRiderBicycle, PlanarPointBicycle and the v0.1
Bicycle (elliptic field).   python tools/bench_models.py [--n 16384] [--steps 50] [--models a,b] [--pair-mode auto|tiled|dense]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cyclistsocialforce_b200 import parameters as P  # noqa: E402
from cyclistsocialforce_b200.engine import AgentGroup, Engine, N_STATES  # noqa: E402
from cyclistsocialforce_b200.synthetic import queues_with_start, synthetic_crowd  # noqa: E402

PARAMS = dict(twod=P.InvPendulumBicycleParameters, invpendulum=P.InvPendulumBicycleParameters,
              balancingrider=P.BalancingRiderBicycleParameters, planarpoint=P.PlanarPointBicycleParameters,
              bicycle=P.BicycleParameters)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--models", default="twod,planarpoint,invpendulum,balancingrider,bicycle")
    ap.add_argument("--pair-mode", default="auto")
    ap.add_argument("--count-pairs", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    for model in a.models.split(","):
        s0, q = synthetic_crowd(a.n, seed=1, n_states=N_STATES[model])
        g = AgentGroup(model, s0, PARAMS[model](), destqueues=list(queues_with_start(s0, q)), dtype=torch.float32,
                       device=dev)
        eng = Engine([g], dtype=torch.float32, device=dev, graph=True, pair_mode=a.pair_mode, count_pairs=a.count_pairs)
        for _ in range(5):
            eng.step()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        for _ in range(a.steps):
            eng.step()
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / a.steps
        # per-agent kernel alone (K2+K3), kernel by kernel
        eng.use_graph = False
        have = eng._pair_and_road()
        torch.cuda.synchronize()
        ev[1].record()
        for _ in range(10):
            eng._agent_step(have)
        ev[2].record()
        torch.cuda.synchronize()
        eng.check_status()
        print(json.dumps({"model": model, "n": a.n, "pair_kernel": "tiled" if eng.tiled else "dense",
                          "evaluated_pair_fraction": (float(eng.pair_stats[0].item()) / max(eng._pair_calls, 1) / a.n ** 2
                                                      if eng.tiled and a.count_pairs else None),
                          "ms_per_step": ms, "agent_steps_per_s": a.n / (ms * 1e-3),
                          "agent_kernel_ms": ev[1].elapsed_time(ev[2]) / 10}), flush=True)


if __name__ == "__main__":
    main()
