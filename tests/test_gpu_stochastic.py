"""Stochastic rider behaviour on the device against the oracle's restatement of
PoleModel.sample_poles (controlbehavior.py:1414-1469) -- the same counter-based random stream on both
sides, so poles, draw counts, gains and the stepped states are compared rider by rider.  Needs a B200."""
import numpy as np
import pytest
import torch

from oracle import csf_oracle as co
from oracle import pole_sampling as ps
from cyclistsocialforce_b200 import parameters as P
from cyclistsocialforce_b200.engine import AgentGroup, Engine
from cyclistsocialforce_b200.synthetic import queues_with_start

pytestmark = pytest.mark.gpu
FILE = "BR1_ImRe5GivenV_pole-model-params.yaml"


def _oracle_params(seed, thresh=0.8333, agent_offset=0):
    return co.default_params("balancingrider", stochastic=True, resample_thresh=thresh, seed=seed,
                             agent_offset=agent_offset, pole_model_file=FILE)


@pytest.mark.parametrize("model_file", [FILE, "BR0_ImRe5GivenV_pole-model-params.yaml"])
def test_first_draw_matches_the_oracle_rider_by_rider(model_file):
    """Construction: every rider draws its first poles at its initial speed (dynamics.py:305-306)."""
    n, seed = 300, 1234
    s0, q = co.synthetic_crowd(n, seed=4, spacing=5.0, n_states=8)
    s0[:, 3] = np.linspace(1.5, 6.9, n)
    g = AgentGroup("balancingrider", s0, P.BalancingRiderBicycleParameters(
        stochastic_control_behavior=True, controlparam_seed=seed, controlparam_filename=model_file),
        destqueues=list(queues_with_start(s0, q)), dtype=torch.float64)
    feats = g.br_poles.cpu().numpy().T
    draws = g.br_draws.cpu().numpy()
    gains = g.br_gains.cpu().numpy().T
    m = ps.load_model(model_file)
    redrawn = 0
    for k in range(n):
        f, used = ps.sample_features(m, s0[k, 3], seed, k)
        assert used == draws[k], (k, used, draws[k])
        assert np.abs(f - feats[k]).max() < 1e-9 * max(1.0, np.abs(f).max()), (k, f, feats[k])
        redrawn += used > 1
        A, B = co.balancingrider_matrices(co.BALANCEASSIST, s0[k, 3])
        ref = co.place_gain(A, B, co.ps_poles(f))
        assert np.abs(gains[k] - ref).max() < 1e-7 * max(1.0, np.abs(ref).max()), (k, gains[k], ref)
    assert np.all(feats[:, [0, 1, 3]] <= 0)
    assert np.all(g.br_vlast.cpu().numpy() == s0[:, 3])
    print("riders that re-drew an out-of-range / unstable sample:", redrawn)


def test_device_sampler_distribution_matches_the_reference():
    """20,000 riders at one speed: the empirical CDF of the device's samples at the reference's quantiles
    (40,000 reference samples, tests/golden/golden_polemodel.npz) within sampling error."""
    import os
    gold = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_polemodel.npz")))
    n = 20000
    qs = gold["qlevels"]
    for v in (2.0, 5.0):
        s0 = np.zeros((n, 8))
        s0[:, 0] = np.arange(n) * 3.0
        s0[:, 3] = v
        g = AgentGroup("balancingrider", s0, P.BalancingRiderBicycleParameters(
            stochastic_control_behavior=True, controlparam_seed=77), dtype=torch.float64)
        feats = g.br_poles.cpu().numpy().T
        refq = gold[f"BR1_v{v}_quantiles"]
        for c in range(5):
            cdf = (feats[:, c][None, :] <= refq[:, c][:, None]).mean(axis=1)
            tol = 4.5 * np.sqrt(qs * (1 - qs) * (1 / n + 1 / 40000)) + 1e-3
            assert np.all(np.abs(cdf - qs) < tol), (v, c, np.abs(cdf - qs).max())


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_stochastic_crowd_steps_and_resampling(dtype):
    """A crowd of stochastic riders that brake and accelerate (speed changes beyond the re-sampling
    threshold): states, draw counts and poles against the oracle, step by step."""
    n, steps, seed, thresh = 24, 60, 5, 0.05
    s0, q = co.synthetic_crowd(n, seed=31, spacing=4.0, n_states=8)
    s0[:, 3] = np.linspace(2.0, 6.0, n)
    A = co.Agents("balancingrider", s0, params=_oracle_params(seed, thresh), v_desired=np.full(n, 5.0))
    for k in range(n):
        A.set_destinations(k, q[k, :, 0], q[k, :, 1])
    W = co.World([A])
    g = AgentGroup("balancingrider", s0, P.BalancingRiderBicycleParameters(
        stochastic_control_behavior=True, controlparam_seed=seed, controlparam_resampling_speedthresh=thresh),
        vd_default=5.0, destqueues=list(queues_with_start(s0, q)), dtype=dtype)
    eng = Engine([g], dtype=dtype)
    assert np.array_equal(g.br_draws.cpu().numpy(), A.br_draws)
    tol = 1e-8 if dtype == torch.float64 else 5e-3
    for k in range(steps):
        W.step()
        eng.step()
        if dtype == torch.float64:
            assert np.array_equal(g.br_draws.cpu().numpy(), A.br_draws), k
            assert np.abs(g.br_poles.cpu().numpy().T - A.br_feats).max() < 1e-8
        s = g.states_numpy()
        assert np.abs(s[:, :2] - A.s[:, :2]).max() < tol, (k, np.abs(s[:, :2] - A.s[:, :2]).max())
        assert np.abs(s[:, 3:] - A.s[:, 3:]).max() < tol * 100, (k, np.abs(s[:, 3:] - A.s[:, 3:]).max())
    assert A.br_draws.min() >= 2                    # every rider re-sampled at least once
    eng.check_status()


def test_random_stream_follows_the_rider_not_the_grouping():
    """The stream is keyed by the rider (``stream_ids``): a crowd split into two groups (or shards) draws the
    poles of the single-group crowd, and the key travels with the rider through churn."""
    n, seed = 64, 9
    s0, q = co.synthetic_crowd(n, seed=2, spacing=4.0, n_states=8)
    par = dict(stochastic_control_behavior=True, controlparam_seed=seed)
    g = AgentGroup("balancingrider", s0, P.BalancingRiderBicycleParameters(**par), dtype=torch.float64)
    ga = AgentGroup("balancingrider", s0[:40], P.BalancingRiderBicycleParameters(**par), dtype=torch.float64)
    gb = AgentGroup("balancingrider", s0[40:], P.BalancingRiderBicycleParameters(**par), dtype=torch.float64,
                    stream_ids=np.arange(40, n))
    both = torch.cat([ga.br_poles, gb.br_poles], dim=1).cpu().numpy()
    assert np.array_equal(both, g.br_poles.cpu().numpy())
    sub = g.select([5, 50, 7])
    assert sub.br_stream.cpu().tolist() == [5, 50, 7]


def test_facade_stochastic_riders():
    """BalancingRiderBicycle(params=BalancingRiderBicycleParameters(stochastic_control_behavior=True)) through
    SocialForceIntersection: every rider has its own poles (own stream), the crowd steps, and churn keeps a
    rider's poles and draw count."""
    from cyclistsocialforce_b200.intersection import SocialForceIntersection
    from cyclistsocialforce_b200.vehicle import BalancingRiderBicycle
    bikes = []
    for k in range(6):
        b = BalancingRiderBicycle((4.0 * k, 0.0, 0.0, 4.0, 0, 0, 0, 0), id=f"r{k}",
                                  params=P.BalancingRiderBicycleParameters(stochastic_control_behavior=True))
        b.setDestinations((4.0 * k + 60.0, 4.0 * k + 120.0), (5.0, -5.0))
        bikes.append(b)
    ins = SocialForceIntersection(bikes, dtype=torch.float64)
    g = ins._groups[0]
    assert g.n == 6                                            # one parameter set -> one device group
    poles = g.br_poles.cpu().numpy().T
    assert len({tuple(np.round(p, 12)) for p in poles}) == 6   # six different samples
    m = ps.load_model(FILE)
    for k, b in enumerate(bikes):
        f, used = ps.sample_features(m, 4.0, 0, b._stream_id)
        assert np.abs(f - poles[k]).max() < 1e-9 and used == int(g.br_draws[k])
    for _ in range(20):
        ins.step()
    before = {b.id: ins._groups[0].br_poles[:, b._k].cpu().numpy().copy() for b in bikes}
    ins.remove_road_user(2)
    g2 = ins._groups[0]
    for b in bikes[:2] + bikes[3:]:
        assert np.array_equal(g2.br_poles[:, b._k].cpu().numpy(), before[b.id])
    for _ in range(5):
        ins.step()
    assert np.all(np.isfinite(np.array([b.s for b in ins.vehicles])))


@pytest.mark.parametrize("kind", ["poles", "gains"])
def test_fixed_poles_and_fixed_gains(kind):
    """BalancingRiderBicycleParameters(poles=...) / (gains=...) (reference parameters.py:1306-1314,
    dynamics.py:602-615): the pole model is ignored; with fixed poles the gains are still re-designed for
    every new speed, with fixed gains never."""
    n, steps = 20, 15
    s0, q = co.synthetic_crowd(n, seed=41, spacing=4.0, n_states=8)
    s0[:, 3] = np.linspace(3.0, 6.0, n)
    poles = np.array([-9.0, -1.1 + 2.0j, -1.1 - 2.0j, -1.7 + 6.5j, -1.7 - 6.5j])
    gains = np.array([-13.1, 1.1, -6.7, -0.11, -11.4])
    if kind == "poles":
        par = P.BalancingRiderBicycleParameters(poles=poles)
        op = co.default_params("balancingrider", fixed_poles=poles)
    else:
        par = P.BalancingRiderBicycleParameters(gains=gains)
        op = co.default_params("balancingrider", fixed_gains=gains)
    A = co.Agents("balancingrider", s0, params=op, v_desired=np.full(n, 5.0))
    for k in range(n):
        A.set_destinations(k, q[k, :, 0], q[k, :, 1])
    W = co.World([A])
    g = AgentGroup("balancingrider", s0, par, vd_default=5.0, destqueues=list(queues_with_start(s0, q)),
                   dtype=torch.float64)
    eng = Engine([g], dtype=torch.float64)
    assert np.abs(g.br_gains.cpu().numpy().T - A.gains).max() < 1e-7 * np.abs(A.gains).max()
    for k in range(steps):
        W.step()
        eng.step()
        assert np.abs(g.states_numpy() - A.s).max() < 1e-8, k
    if kind == "gains":
        assert np.array_equal(g.br_gains.cpu().numpy().T, np.tile(gains, (n, 1)))
    eng.check_status()
