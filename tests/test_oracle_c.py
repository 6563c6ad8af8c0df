"""The C/OpenMP restatement (bench.py's CPU baseline) agrees with the numpy oracle, which is
pinned to the reference.  CPU only."""
import numpy as np

from oracle import c_port, csf_oracle as co
from helpers import oracle_world


def test_c_pair_forces_match_numpy_oracle():
    s0, q = co.synthetic_crowd(300, seed=12, spacing=2.5)
    p = co.default_params("twod")
    fp = co.field_params_array([p])[0]
    for p2r in (False, True):
        ref = co.pair_forces(s0[:, 0], s0[:, 1], s0[:, 2], fp, p2r=p2r)
        got = c_port.pair_forces(s0[:, 0], s0[:, 1], s0[:, 2], fp, p2r=p2r)
        assert np.abs(got - ref).max() < 1e-12
    sub = np.array([5, 17, 299], dtype=np.int64)
    assert np.abs(c_port.pair_forces(s0[:, 0], s0[:, 1], s0[:, 2], fp, tgt=sub) - ref_sub(s0, fp, sub)).max() < 1e-12


def ref_sub(s0, fp, sub):
    return co.pair_forces(s0[:, 0], s0[:, 1], s0[:, 2], fp, tgt=sub)


def test_c_twod_steps_match_numpy_oracle():
    n = 48
    s0, q = co.synthetic_crowd(n, seed=21, spacing=3.0)
    W = oracle_world("twod", s0, np.full(n, 5.0), q)
    Cc = c_port.TwoDCrowdC(s0, q)
    worst = 0.0
    for _ in range(40):
        W.step()
        Cc.step()
        worst = max(worst, np.abs(W.groups[0].s - Cc.s).max(), np.abs(W.groups[0].force - Cc.force).max())
    assert worst < 1e-10
    assert np.array_equal(W.groups[0].ptr, Cc.ptr)


def test_c_demo_geometry_golden(golden):
    """Last-destination spline branch + nav machine of the C port, against the reference vectors."""
    g = golden
    Cc = c_port.TwoDCrowdC(g["demo_s0"], g["demo_dests"], vd=g["demo_vd"])
    keep = g["demo_twod_steps"].tolist()
    S = []
    for k in range(1, 1501):
        Cc.step()
        if k in keep:
            S.append(Cc.s.copy())
    assert np.abs(np.array(S) - g["demo_twod_s"]).max() < 5e-11
    assert np.array_equal(Cc.ptr, g["demo_twod_ptr"])
