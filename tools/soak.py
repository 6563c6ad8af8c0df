#!/usr/bin/env python
"""Soak run (needs a B200): the benchmark crowd stepped for a few thousand steps through the CUDA-graph path --
status words polled every step, re-sorts every 64 steps, trajectory stream on -- then compared on a sample of
agents with the f64 build run over the same steps (drift report), and a churn loop on a smaller crowd.

    python tools/soak.py [N] [steps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cyclistsocialforce_b200 import parameters as P  # noqa: E402
from cyclistsocialforce_b200.engine import AgentGroup, Engine  # noqa: E402
from cyclistsocialforce_b200.synthetic import queues_with_start, spatial_order, synthetic_crowd  # noqa: E402
from cyclistsocialforce_b200.trajstream import TrajectoryStream  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
s0, q = synthetic_crowd(n, seed=1)
o = spatial_order(s0[:, 0], s0[:, 1])
s0, q = s0[o], q[o]
origin, extent = P.payload_frame([s0[:, :2], q[..., :2]])
res = {}
for dtype in (torch.float32, torch.float64):
    g = AgentGroup("twod", s0, P.InvPendulumBicycleParameters(), destqueues=list(queues_with_start(s0, q)), dtype=dtype)
    eng = Engine([g], dtype=dtype, extent=extent, origin=origin, graph=True)
    ts = TrajectoryStream(eng, chunk_steps=64) if dtype == torch.float32 else None
    k = steps if dtype == torch.float32 else min(steps, 300)
    t0 = time.perf_counter()
    for i in range(k):
        eng.step()
        if ts is not None and i % 8 == 0:
            ts.append()
        if i == 299:
            res[(dtype, 300)] = g.states_numpy()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    eng.check_status()
    s = g.states_numpy()
    assert np.all(np.isfinite(s)), "non-finite state"
    print(f"{dtype}: {k} steps of {n} agents in {dt:.2f} s wall ({n * k / dt / 1e6:.1f} M agent-steps/s incl. host), "
          f"speed range {s[:, 3].min():.2f}..{s[:, 3].max():.2f} m/s, extent {np.ptp(s[:, 0]):.0f} x {np.ptp(s[:, 1]):.0f} m")
    if ts is not None:
        rec = ts.drain()
        print(f"   trajectory stream: {sum(c['steps'] for c in rec)} recorded steps in {len(rec)} chunks")
d = np.abs(res[(torch.float32, 300)] - res[(torch.float64, 300)])
print("f32 vs f64 after 300 steps: median |diff| %.2e, 99th percentile %.2e, max %.2e (positions m, angles rad)"
      % (np.median(d.max(axis=1)), np.percentile(d.max(axis=1), 99), d.max()))
print("soak ok")
