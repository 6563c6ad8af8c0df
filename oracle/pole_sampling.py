"""Oracle (test infrastructure) for the stochastic rider behaviour of BalancingRiderBicycle:
``BalancingRiderBicycleParameters.update_control_params`` in stochastic mode (reference
src/cyclistsocialforce/parameters.py:1376-1402) -> ``PoleModel.sample_poles``
(controlbehavior.py:1414-1469) -> ``PoleModel.sample`` (:1337-1412) ->
``ConditionalGaussianMixture._get_conditional_gmm`` (:477-533) and the pre-processing pipeline
(``PreprocessingPipeline.transform / inverse_transform`` :913-985, ``LogTransformer`` :613-695, sklearn's
``PowerTransformer(method="yeo-johnson", standardize=True)``).

Restated in numpy, every deterministic stage pinned against the reference's own objects through
tests/golden/golden_polemodel.npz (tests/golden/make_polemodels.py).  The random stage is restated with a
counter-based generator (Philox-4x32-10) so that the device sampler can be checked draw for draw; that the
resulting distribution is the reference's is checked against 40,000 reference samples per speed.
"""
import json
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cyclistsocialforce_b200", "data",
                     "pole_models.json")
_EPS = np.spacing(1.0)


def load_model(filename="BR1_ImRe5GivenV_pole-model-params.yaml"):
    with open(_DATA) as f:
        m = json.load(f)[filename]
    return {k: (np.array(v) if isinstance(v, list) and k not in ("features",) else v) for k, v in m.items()}


# ---- Yeo-Johnson (sklearn.preprocessing.PowerTransformer) ------------------------------------------------
def yeo_johnson(x, lam):
    x = np.asarray(x, float)
    out = np.empty_like(x)
    pos = x >= 0
    if abs(lam) < _EPS:
        out[pos] = np.log1p(x[pos])
    else:
        out[pos] = (np.power(x[pos] + 1, lam) - 1) / lam
    if abs(lam - 2) > _EPS:
        out[~pos] = -(np.power(-x[~pos] + 1, 2 - lam) - 1) / (2 - lam)
    else:
        out[~pos] = -np.log1p(-x[~pos])
    return out


def yeo_johnson_inverse(y, lam):
    """NaN where y is outside the range of the transform (the reference resamples those)."""
    y = np.asarray(y, float)
    out = np.empty_like(y)
    pos = y >= 0
    with np.errstate(invalid="ignore"):
        if abs(lam) < _EPS:
            out[pos] = np.exp(y[pos]) - 1
        else:
            out[pos] = np.power(y[pos] * lam + 1, 1 / lam) - 1
        if abs(lam - 2) > _EPS:
            out[~pos] = 1 - np.power(-(2 - lam) * y[~pos] + 1, 1 / (2 - lam))
        else:
            out[~pos] = 1 - np.exp(-y[~pos])
    return out


# ---- pipeline -----------------------------------------------------------------------------------------
def transform_given(m, v):
    """The conditioning speed in the model's feature space (PoleModel.sample :1358-1365: the speed is not a
    log-shifted feature, so only the power transform and the scaler act on it)."""
    ig = int(m["index_given"])
    return (yeo_johnson(np.atleast_1d(float(v)), m["lambdas"][ig]) - m["scaler_mean"][ig]) / m["scaler_scale"][ig]


def inverse_transform(m, z):
    """Transformed pole features (n, 5) -> pole features [p0_real, p1_real, p1_imag, p2_real, p2_imag]:
    un-standardise, inverse Yeo-Johnson, then x = sign * (exp(y) + a) on the log-shifted (real-part) features
    (PreprocessingPipeline.inverse_transform :962-985)."""
    ig = int(m["index_given"])
    idx = [i for i in range(len(m["features"])) if i != ig]
    z = np.atleast_2d(np.asarray(z, float))
    out = np.empty_like(z)
    logf = list(m["log_features"])
    for c, i in enumerate(idx):
        y = yeo_johnson_inverse(z[:, c] * m["scaler_scale"][i] + m["scaler_mean"][i], m["lambdas"][i])
        if i in logf:
            j = logf.index(i)
            y = (np.exp(y) + m["log_a"][j]) / m["log_sign"][j]
        out[:, c] = y
    return out


def conditional_gmm(m, xt):
    """Weights, means, covariances of the mixture conditioned on the transformed speed xt
    (ConditionalGaussianMixture._get_conditional_gmm :477-533)."""
    ig = int(m["index_given"])
    idx = [i for i in range(len(m["features"])) if i != ig]
    xt = float(np.ravel(xt)[0])
    w, mu, cov = [], [], []
    for k in range(len(m["weights"])):
        S, mk = m["covariances"][k], m["means"][k]
        var_g = S[ig, ig]
        c = S[idx, ig]
        mu.append(mk[idx] + c / var_g * (xt - mk[ig]))
        cov.append(S[np.ix_(idx, idx)] - np.outer(c, c) / var_g)
        w.append(m["weights"][k] * np.exp(-0.5 * (xt - mk[ig]) ** 2 / var_g) / np.sqrt(2 * np.pi * var_g))
    w = np.array(w) / np.sum(w)
    if np.any(w == 0.0):
        w[w == 0.0] = np.finfo(float).eps * len(w)
        w = w / w.sum()
    return w, np.array(mu), np.array(cov)


def features_to_poles(f):
    """polefeaturetable_to_polearray(features='ImRe') (controlbehavior.py:65-89)."""
    f = np.atleast_2d(f)
    p1, p2 = f[:, 1] + 1j * f[:, 2], f[:, 3] + 1j * f[:, 4]
    return np.stack([f[:, 0] + 0j, p1, np.conj(p1), p2, np.conj(p2)], axis=1)


# ---- counter-based random numbers (Philox-4x32-10, Salmon et al. 2011) -----------------------------------
_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32(counter, key):
    """counter: 4 uint32, key: 2 uint32 -> 4 uint32 (10 rounds)."""
    c = [int(x) & 0xFFFFFFFF for x in counter]
    k = [int(x) & 0xFFFFFFFF for x in key]
    for _ in range(10):
        p0, p1 = _M0 * c[0], _M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + _W0) & 0xFFFFFFFF, (k[1] + _W1) & 0xFFFFFFFF]
    return c


def draw(seed, agent, index):
    """One uniform in (0, 1) and five standard normals for draw ``index`` of road user ``agent``: two
    Philox blocks, Box-Muller on consecutive words."""
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    r = philox4x32((agent & 0xFFFFFFFF, index, 0, (agent >> 32) & 0xFFFFFFFF), key) + \
        philox4x32((agent & 0xFFFFFFFF, index, 1, (agent >> 32) & 0xFFFFFFFF), key)
    u = [(x + 0.5) * 2.0 ** -32 for x in r]
    n = []
    for a, b in ((1, 2), (3, 4), (5, 6)):
        rad = np.sqrt(-2.0 * np.log(u[a]))
        n += [rad * np.cos(2 * np.pi * u[b]), rad * np.sin(2 * np.pi * u[b])]
    return u[0], np.array(n[:5])


def sample_features(m, v, seed, agent, first_index=0, max_draws=1000):
    """Pole features for road user ``agent`` at speed v: draws first_index, first_index + 1, ... until the
    sample is inside the range of the inverse transform and stable (the reference re-draws such samples,
    PoleModel.sample :1377-1394, sample_poles :1456-1467).  Returns (features, draws used)."""
    w, mu, cov = conditional_gmm(m, transform_given(m, v))
    L = np.linalg.cholesky(cov)
    cw = np.cumsum(w)
    for t in range(max_draws):
        u, z = draw(seed, agent, first_index + t)
        k = int(min(np.searchsorted(cw, u, side="right"), len(w) - 1))
        f = inverse_transform(m, (mu[k] + L[k] @ z)[None, :])[0]
        if np.all(np.isfinite(f)) and f[0] <= 0 and f[1] <= 0 and f[3] <= 0:
            return f, t + 1
    raise TimeoutError("no valid pole sample")
