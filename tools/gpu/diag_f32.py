"""Diagnose the largest f32 per-step force errors: which term (repulsive sum, destination force) carries them."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import csf_oracle as co
from gpu_helpers import make_engine
from cyclistsocialforce_b200.engine import _FIELD_AXIS, _STATE_COLS

model = sys.argv[1] if len(sys.argv) > 1 else "twod"
n, steps = int(sys.argv[2]) if len(sys.argv) > 2 else 768, int(sys.argv[3]) if len(sys.argv) > 3 else 16
s0, q = co.synthetic_crowd(n, seed=33, spacing=3.0, n_states=8)
e64, g64 = make_engine(model, s0, 5.0, q, dtype=torch.float64)
e32, g32 = make_engine(model, s0, 5.0, q, dtype=torch.float32, q_scale=e64.q_scale)
fields = [f for f in _STATE_COLS + tuple(_FIELD_AXIS) if f not in ("destq", "dest_len", "vd_default") and getattr(g64, f) is not None]
for step in range(steps):
    for name in fields:
        getattr(g32, name).copy_(getattr(g64, name).to(getattr(g32, name).dtype))
    e32.pack()
    # forces only (K1 + K2), both builds, then restore the nav state of the f64 build by stepping a copy
    snap = {name: getattr(g64, name).clone() for name in fields}
    e64.calc_forces(); e32.calc_forces()
    F64, F32 = e64.force.cpu().numpy(), e32.force.cpu().numpy().astype(float)
    R64, R32 = e64.frep.cpu().numpy(), e32.frep.cpu().numpy().astype(float)
    D64, D32 = e64.fdest.cpu().numpy(), e32.fdest.cpu().numpy().astype(float)
    for name in fields:
        getattr(g64, name).copy_(snap[name])
        getattr(g32, name).copy_(snap[name].to(getattr(g32, name).dtype))
    ef = np.linalg.norm(F32 - F64, axis=1) / np.maximum(np.linalg.norm(F64, axis=1), 1.0)
    k = int(np.argmax(ef))
    if ef[k] > 3e-5:
        print(f"step {step} agent {k} ef={ef[k]:.3e} |F|={np.linalg.norm(F64[k]):.4f} F64={F64[k]} dF={F32[k]-F64[k]}\n"
              f"   frep64={R64[k]} dfrep={R32[k]-R64[k]} |frep|={np.linalg.norm(R64[k]):.4f}\n"
              f"   fdest64={D64[k]} dfdest={D32[k]-D64[k]} ptr={int(g64.dest_ptr[k])} v={float(g64.v[k]):.4f}", flush=True)
    e64.step(); e32.step()
print("done")
