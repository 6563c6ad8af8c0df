"""Live check of the restated oracle against the reference's own code (needs
/root/reference, i.e. the build container; skipped on the GPU box)."""
import numpy as np
import pytest

from oracle import csf_oracle as co
from oracle import ref_harness as rh
from helpers import oracle_world

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not rh.reference_available(), reason="no /root/reference")]


def _ref_world(cls, s0, vd, dests, **kw):
    bikes = []
    for k in range(s0.shape[0]):
        b = cls(tuple(s0[k, :cls.N_STATES]), id=str(k))
        b.params.v_desired_default = float(vd[k])
        b.setDestinations(dests[k][:, 0], dests[k][:, 1], stop=dests[k][:, 2])
        bikes.append(b)
    return rh.headless_intersection(bikes, **kw)


@pytest.mark.parametrize("model,clsname,n,steps", [
    ("twod", "TwoDBicycle", 16, 120),
    ("planarpoint", "PlanarPointBicycle", 6, 60),
    ("invpendulum", "InvPendulumBicycle", 8, 80),
    ("bicycle", "Bicycle", 12, 80),
])
def test_seeded_crowd(model, clsname, n, steps):
    veh = rh.modules()[0]
    s0, q = co.synthetic_crowd(n, seed=11, spacing=3.0, n_states=8)
    vd = np.full(n, 5.0)
    ins = _ref_world(getattr(veh, clsname), s0, vd, q)
    W = oracle_world(model, s0, vd, q)
    ms = mf = 0.0
    for _ in range(steps):
        ins.step()
        W.step()
        sr = np.array([v.s for v in ins.vehicles])
        fr = np.array([v.force for v in ins.vehicles])
        ms = max(ms, np.abs(sr - W.groups[0].s).max())
        mf = max(mf, np.abs(fr - W.groups[0].force).max())
    assert ms < 1e-10 and mf < 1e-10, (ms, mf)
    assert [v.destpointer for v in ins.vehicles] == W.groups[0].ptr.tolist()


def test_mask_matches_get_untracked_foes():
    veh = rh.modules()[0]
    s0, q = co.synthetic_crowd(20, seed=5, spacing=2.0, n_states=8)
    for rule in ("unregulated", "p2r"):
        ins = _ref_world(veh.TwoDBicycle, s0, np.full(20, 5.0), q, priority_rule=rule)
        tr = co.tracked_mask(s0[:, 0], s0[:, 1], s0[:, 2], 2 * np.pi / 3, p2r=rule == "p2r")
        assert np.array_equal(tr, ~ins.get_untracked_foes())
