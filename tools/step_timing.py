"""Diagnostic: pair-phase time and executed pair evaluations as the crowd evolves
(CSF_ST_STEPS: simulate this many steps first, then time the pair phase on the frozen state)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cyclistsocialforce_b200 import parameters as P
from cyclistsocialforce_b200.engine import AgentGroup, Engine
from cyclistsocialforce_b200.synthetic import queues_with_start, spatial_order, synthetic_crowd
N = int(os.environ.get("CSF_BENCH_N", 65536))
K = int(os.environ.get("CSF_ST_STEPS", 200))
s0, q = synthetic_crowd(N, seed=1); o = spatial_order(s0[:, 0], s0[:, 1]); s0, q = s0[o], q[o]
extent = 2.0 * float(max(np.abs(s0[:, :2]).max(), np.abs(q[..., :2]).max())) + 1000.0
g = AgentGroup("twod", s0, P.InvPendulumBicycleParameters(), destqueues=list(queues_with_start(s0, q)), dtype=torch.float32)
eng = Engine([g], dtype=torch.float32, extent=extent, pair_mode="tiled", count_pairs=True)
for step in range(K):
    eng.step()
torch.cuda.synchronize()
eng.resort_every = 10 ** 9
eng.pair_stats.zero_()
eng._pair_and_road()
torch.cuda.synchronize()
ev_pairs = eng.pair_stats.item() / N
eng.pair_stats = None
ts = []
for _ in range(20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng._pair_and_road(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(f"N {N} after {K} steps: pair phase mean {np.mean(ts):.4f} best {np.min(ts):.4f} ms, evaluated/target {ev_pairs:.0f}  "
      f"[TPW={os.environ.get('CSF_TILED_TPW','-')} GROUPS={os.environ.get('CSF_TILED_GROUPS','-')} IPS={os.environ.get('CSF_TILED_ITEMS_PER_SLOT','-')}]", flush=True)
if os.environ.get("CSF_ST_DIAG"):
    st = g.states_numpy()
    perm = eng._tgt_perm.cpu().numpy()
    xs, ys = st[perm, 0], st[perm, 1]
    for B, name in ((64, "tile/block of 64"), (1024, "chunk of 1024")):
        nb = N // B
        X, Y = xs[:nb * B].reshape(nb, B), ys[:nb * B].reshape(nb, B)
        R = np.hypot(X.max(1) - X.min(1), Y.max(1) - Y.min(1)) / 2
        print(name, "R: median %.1f  p90 %.1f  p99 %.1f  max %.1f" % (np.median(R), np.percentile(R, 90), np.percentile(R, 99), R.max()))
    # local density: neighbours within 20 m for a sample
    idx = np.random.default_rng(0).choice(N, 2000, replace=False)
    from scipy.spatial import cKDTree
    tr = cKDTree(st[:, :2])
    cnt = np.array([len(tr.query_ball_point(st[i, :2], 20.0)) for i in idx])
    print("neighbours within 20 m: mean %.1f p99 %.1f max %d (uniform expectation %.1f)" % (cnt.mean(), np.percentile(cnt, 99), cnt.max(), np.pi * 400 / 16))
    print("x range", st[:, 0].min(), st[:, 0].max(), "y range", st[:, 1].min(), st[:, 1].max(), "finite", np.isfinite(st).all())
