"""Stochastic rider behaviour: every deterministic stage of the reference's
PoleModel.sample_poles against answers produced by the reference's own objects
(tests/golden/golden_polemodel.npz, generator: tests/golden/make_polemodels.py), and the distribution of
the restated sampler against 40,000 reference samples per speed."""
import numpy as np
import pytest

from oracle import pole_sampling as ps

FILES = {"BR1": "BR1_ImRe5GivenV_pole-model-params.yaml", "BR0": "BR0_ImRe5GivenV_pole-model-params.yaml"}


@pytest.fixture(scope="module")
def gold():
    import os
    return dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_polemodel.npz")))


@pytest.mark.parametrize("key", list(FILES))
def test_deterministic_stages_match_the_reference(gold, key):
    m = ps.load_model(FILES[key])
    inv = ps.inverse_transform(m, gold[f"{key}_z"])
    ref = gold[f"{key}_z_inverse"]
    assert np.array_equal(np.isnan(inv), np.isnan(ref))            # the same samples fall outside the transform's range
    ok = ~np.isnan(ref)
    assert np.abs(inv[ok] - ref[ok]).max() < 1e-9 * max(1.0, np.abs(ref[ok]).max())
    for v in gold["speeds"]:
        tag = f"{key}_v{v}"
        xt = ps.transform_given(m, v)
        assert np.abs(xt - gold[tag + "_xt"]).max() < 1e-12
        w, mu, cov = ps.conditional_gmm(m, xt)
        assert np.abs(w - gold[tag + "_w"]).max() < 1e-12
        assert np.abs(mu - gold[tag + "_mu"]).max() < 1e-12
        assert np.abs(cov - gold[tag + "_cov"]).max() < 1e-12


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox-4x32-10."""
    assert ps.philox4x32((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert ps.philox4x32((0xffffffff,) * 4, (0xffffffff,) * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert ps.philox4x32((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_poles_from_features():
    p = ps.features_to_poles(np.array([[-3.0, -1.0, 2.0, -4.0, 7.0]]))
    assert np.array_equal(p[0], [-3 + 0j, -1 + 2j, -1 - 2j, -4 + 7j, -4 - 7j])


@pytest.mark.parametrize("key", ["BR1"])
def test_sampler_distribution_matches_the_reference(gold, key):
    """The counter-based sampler draws from the reference's distribution: at every reference quantile level
    the empirical CDF of 6,000 restated samples is within sampling error (4.5 sigma of a binomial + the
    reference's own error), per feature and speed; means agree within 5 standard errors."""
    m = ps.load_model(FILES[key])
    q = gold["qlevels"]
    n = 6000
    for v in (2.0, 5.0):
        tag = f"{key}_v{v}"
        feats = np.array([ps.sample_features(m, v, seed=99, agent=a)[0] for a in range(n)])
        refq = gold[tag + "_quantiles"]
        for c in range(5):
            cdf = (feats[:, c][None, :] <= refq[:, c][:, None]).mean(axis=1)
            tol = 4.5 * np.sqrt(q * (1 - q) * (1 / n + 1 / 40000)) + 1e-3
            assert np.all(np.abs(cdf - q) < tol), (v, c, np.abs(cdf - q).max())
        se = np.sqrt(np.diag(gold[tag + "_cov_samples"]) * (1 / n + 1 / 40000))
        assert np.all(np.abs(feats.mean(axis=0) - gold[tag + "_mean"]) < 5 * se)
        assert np.all(feats[:, [0, 1, 3]] < 0)                      # stable poles


@pytest.mark.parametrize("key", list(FILES))
def test_device_parameter_block_holds_the_conditional_mixture(key):
    """What the host hands to the device sampler (CsfAgentParams.br_*: per component the speed's mean and
    variance, the pole features' mean, slope against the speed and the Cholesky factor of the Schur
    complement) reproduces the oracle's conditional mixture -- hence the reference's -- at any speed."""
    from cyclistsocialforce_b200 import parameters as P
    par = P.BalancingRiderBicycleParameters(stochastic_control_behavior=True, controlparam_filename=FILES[key],
                                            controlparam_seed=(7 << 40) + 3, controlparam_resampling_speedthresh=0.5)
    c = par.to_agent_params(1.0, 8, 128)
    m = ps.load_model(FILES[key])
    assert c.br_stochastic == 1 and c.br_n_comp == len(m["weights"]) and c.br_seed == (7 << 40) + 3
    assert c.br_resample_thresh == 0.5
    for v in (1.0, 3.3, 6.9):
        xt = float(ps.transform_given(m, v)[0])
        w, mu, cov = ps.conditional_gmm(m, xt)
        ww = np.array([c.br_w[k] * np.exp(-0.5 * (xt - c.br_mu_g[k]) ** 2 / c.br_var_g[k]) / np.sqrt(2 * np.pi * c.br_var_g[k])
                       for k in range(c.br_n_comp)])
        assert np.abs(ww / ww.sum() - w).max() < 1e-13
        for k in range(c.br_n_comp):
            mk = np.array([c.br_mu[k][i] + c.br_slope[k][i] * (xt - c.br_mu_g[k]) for i in range(5)])
            assert np.abs(mk - mu[k]).max() < 1e-12
            L = np.zeros((5, 5))
            for i in range(5):
                for j in range(i + 1):
                    L[i, j] = c.br_chol[k][i * (i + 1) // 2 + j]
            assert np.abs(L @ L.T - cov[k]).max() < 1e-12


def test_unknown_pole_model_and_component_raise_like_the_reference():
    from cyclistsocialforce_b200 import parameters as P
    with pytest.raises(FileNotFoundError):
        P.BalancingRiderBicycleParameters(stochastic_control_behavior=True, controlparam_filename="nope.yaml")
    with pytest.raises(ValueError):                                      # parameters.py:1369-1373
        P.BalancingRiderBicycleParameters(stochastic_control_behavior=True, controlparam_polemodel_component=7)


def test_oracle_agents_resample_on_speed_change():
    """The oracle's stochastic BalancingRider: poles are re-drawn exactly when the speed has moved by more
    than the threshold since the last draw (parameters.py:1398-1402)."""
    from oracle import csf_oracle as co
    p = co.default_params("balancingrider", stochastic=True, resample_thresh=0.5, seed=3, agent_offset=0,
                          pole_model_file=FILES["BR1"])
    A = co.Agents("balancingrider", np.array([[0, 0, 0, 4.0, 0, 0, 0, 0.0]]), params=p)
    d0, f0 = int(A.br_draws[0]), A.br_feats[0].copy()
    A._br_gains(4.4, 0)
    assert A.br_draws[0] == d0 and np.array_equal(A.br_feats[0], f0) and A.br_vlast[0] == 4.0
    A._br_gains(4.6, 0)
    assert A.br_draws[0] > d0 and not np.array_equal(A.br_feats[0], f0) and A.br_vlast[0] == 4.6


def test_fixed_poles_and_gains_parameters():
    """poles= / gains= (reference parameters.py:1306-1314): features of the fixed poles, flag for fixed gains."""
    from cyclistsocialforce_b200 import parameters as P
    par = P.BalancingRiderBicycleParameters(poles=[-1.7 - 6.5j, -9.0, -1.1 + 2.0j, -1.7 + 6.5j, -1.1 - 2.0j])
    assert par.controlparam_fix and not par.stochastic_control_behavior
    assert par.pole_intercept == (-9.0, -1.1, 2.0, -1.7, 6.5) and par.pole_slope == (0.0,) * 5
    assert [complex(z) for z in par.poles_at(3.3)] == [-9 + 0j, -1.1 + 2j, -1.1 - 2j, -1.7 + 6.5j, -1.7 - 6.5j]
    c = par.to_agent_params(1.0, 8, 128)
    assert c.br_fixed_gains == 0 and c.br_stochastic == 0 and list(c.br_pole_coef) == [0.0] * 5
    g = P.BalancingRiderBicycleParameters(gains=(-13.0, 1.0, -6.0, -0.1, -11.0), stochastic_control_behavior=True)
    assert g.to_agent_params(1.0, 8, 128).br_fixed_gains == 1 and not g.stochastic_control_behavior
    with pytest.raises(ValueError):
        P.BalancingRiderBicycleParameters(poles=[-1, -2 + 1j, -2 - 1j, -3 + 1j, -3 - 2j])
    with pytest.raises(ValueError):
        P.BalancingRiderBicycleParameters(gains=(1.0, 2.0))
