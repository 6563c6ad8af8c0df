"""Where the warps of the tiled pair kernel spend their cycles (needs a library built with
CSF_BUILD_DEFINES=-DCSF_TILED_PROF):  python tools/k1_roles.py [N] [steps]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cyclistsocialforce_b200 import parameters as P
from cyclistsocialforce_b200.engine import AgentGroup, Engine
from cyclistsocialforce_b200.synthetic import queues_with_start, spatial_order, synthetic_crowd

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
s0, q = synthetic_crowd(n, seed=1)
o = spatial_order(s0[:, 0], s0[:, 1]); s0, q = s0[o], q[o]
g = AgentGroup("twod", s0, P.InvPendulumBicycleParameters(), destqueues=list(queues_with_start(s0, q)), dtype=torch.float32)
eng = Engine([g], dtype=torch.float32, pair_mode="tiled", count_pairs=True)
for _ in range(steps):
    eng.step()
torch.cuda.synchronize()
eng.pair_stats.zero_()
eng._pair_and_road()
torch.cuda.synchronize()
s = eng.pair_stats.cpu().numpy().astype(float)
print(f"N={n}: evaluated pairs {s[0]:.4g} ({s[0] / n:.0f} per target)")
if s[1] > 0:
    print(f"evaluate warps: waiting for a buffer {100 * s[2] / s[1]:.1f}% of their time")
    print(f"filter warps  : waiting for a stage {100 * s[4] / s[3]:.1f}%, for a free slot {100 * s[5] / s[3]:.1f}%")
    print(f"producer warps: waiting for a stage slot {100 * s[7] / s[6]:.1f}%")
    print(f"buffers {s[8]:.0f}, stages {s[9]:.0f}, (target, buffer) units with work {s[10]:.0f} "
          f"({s[0] / max(s[10], 1) / 64:.2f} tiles per unit)")
else:
    print("(library built without -DCSF_TILED_PROF: no cycle accounting)")
