"""Trajectory stream and SUMO pose batch: the device ring + chunked device->host copies must
reproduce the per-step history the reference keeps in vehicle.traj / trajF / F (vehicle.py:320-325,
:1407-1413, intersection.py:860-862), and the batched position hand-over the reference's per-vehicle
traci.vehicle.moveToXY loop (intersection.py:679-688, utils.py:119-121).  Needs a B200."""
import numpy as np
import pytest
import torch

from oracle import csf_oracle as co
from cyclistsocialforce_b200 import parameters as P
from cyclistsocialforce_b200.engine import AgentGroup, Engine
from cyclistsocialforce_b200.intersection import SocialForceIntersection
from cyclistsocialforce_b200.synthetic import queues_with_start
from cyclistsocialforce_b200.trajstream import TrajectoryStream, sumo_poses
from cyclistsocialforce_b200.vehicle import InvPendulumBicycle, TwoDBicycle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_stream_equals_per_step_snapshots(dtype):
    """Two model groups, chunks of 7 steps, 40 steps without draining (both halves of the ring are reused
    several times), then a partial chunk: every recorded step equals the state read back right after it."""
    n = 300
    s0, q = co.synthetic_crowd(n, seed=12, spacing=3.0, n_states=6)
    ga = AgentGroup("twod", s0[:200], P.InvPendulumBicycleParameters(), destqueues=list(queues_with_start(s0[:200], q[:200])), dtype=dtype)
    gb = AgentGroup("invpendulum", s0[200:], P.InvPendulumBicycleParameters(), destqueues=list(queues_with_start(s0[200:], q[200:])), dtype=dtype)
    eng = Engine([ga, gb], dtype=dtype)
    ts = TrajectoryStream(eng, chunk_steps=7)
    snaps_a, snaps_b, snaps_f = [], [], []
    for k in range(40):
        eng.step()
        ts.append()
        snaps_a.append(ga.states_numpy())
        snaps_b.append(gb.states_numpy())
        snaps_f.append(eng.force.cpu().numpy().copy())
    chunks = ts.drain()
    assert sum(c["steps"] for c in chunks) == 40 and ts.launches == 40
    row = 0
    for c in chunks:
        for i in range(c["steps"]):
            for cols, snap, names in ((c["groups"][0], snaps_a[row], ("x", "y", "psi", "v", "delta")),
                                      (c["groups"][1], snaps_b[row], ("x", "y", "psi", "v", "delta", "theta"))):
                got = np.stack([cols[nm][i].astype(float) for nm in names], axis=1)
                assert np.array_equal(got, snap)
            assert np.array_equal(c["force"][i], snaps_f[row])
            row += 1
    assert ts.drain() == []


def test_large_crowd_histories_through_the_facade():
    """100 road users (> 64: the per-step host copy is off) with record_traj=True: vehicle.traj, trajF and F
    filled from the stream equal the per-step values."""
    n, steps = 100, 45
    s0, q = co.synthetic_crowd(n, seed=8, spacing=3.0)
    bikes = []
    for k in range(n):
        b = TwoDBicycle(tuple(s0[k]), id=f"b{k}", saveForces=True)
        b.setDestinations(q[k, :, 0], q[k, :, 1])
        bikes.append(b)
    ins = SocialForceIntersection(bikes, dtype=torch.float64, record_traj=True, traj_chunk_steps=16)
    assert ins._traj_stream is not None
    states, forces = [], []
    for _ in range(steps):
        ins.step()
        states.append(ins._groups[0].states_numpy())
        forces.append(ins._engine.force.cpu().numpy().copy())
    for k in (0, 17, 99):
        tr = bikes[k].traj                                # flushes the stream
        assert np.array_equal(tr[:, 0], s0[k][:5])
        for t in range(steps):
            assert np.array_equal(tr[:, t + 1], states[t][k])
            assert np.array_equal(bikes[k].trajF[:, t + 1], forces[t][k])
        assert np.allclose(bikes[k].F, [np.hypot(*forces[t][k]) for t in range(steps)], rtol=0, atol=0)
    ins.remove_road_user(3)                               # churn re-binds the stream; histories stay
    for _ in range(5):
        ins.step()
    assert np.array_equal(bikes[17].traj[:, steps], states[-1][17])
    assert np.all(np.isfinite(bikes[17].traj[:, steps + 5]))
    assert np.abs(bikes[17].traj[:, steps + 5] - bikes[17].s).max() == 0.0


class _FakeTraci:
    class _Vehicle:
        def __init__(self):
            self.calls = []

        def moveToXY(self, vid, edge, lane, x, y, angle=None, keepRoute=None):
            self.calls.append((vid, edge, lane, x, y, angle, keepRoute))

    def __init__(self):
        self.vehicle = self._Vehicle()


def test_batched_sumo_position_update():
    """activate_sumo_cosimulation with a traci-like client: after every step each road user is moved to its
    new position with the reference's arguments (intersection.py:679-688) and SUMO angle convention."""
    n = 40
    s0, q = co.synthetic_crowd(n, seed=5, spacing=3.0, n_states=6)
    bikes = []
    for k in range(n):
        cls = TwoDBicycle if k % 2 else InvPendulumBicycle
        b = cls(tuple(s0[k][:cls.N_STATES]), id=f"v{k}")
        b.setDestinations(q[k, :, 0], q[k, :, 1])
        bikes.append(b)
    client = _FakeTraci()
    ins = SocialForceIntersection(bikes, dtype=torch.float64, activate_sumo_cosimulation=True, sumo_client=client)
    for _ in range(3):
        ins.step()
    calls = client.vehicle.calls
    assert len(calls) == 3 * n
    last = {c[0]: c for c in calls[-n:]}
    for b in bikes:
        vid, edge, lane, x, y, ang, keep = last[b.id]
        s = b.s
        th = np.pi / 2 - s[2]
        ref = 360.0 * (th + (2 * np.pi if th < 0 else 0.0)) / (2 * np.pi)       # utils.py:119-121, :142-148
        assert (edge, lane, keep) == ("", -1, 6)
        assert x == s[0] and y == s[1]
        assert abs(ang - ref) < 1e-9 or abs(abs(ang - ref) - 360.0) < 1e-9
    g = ins._groups[0]
    poses = sumo_poses(g)
    assert poses.shape == (g.n, 3) and np.all((poses[:, 2] >= 0) & (poses[:, 2] < 360.0 + 1e-9))
