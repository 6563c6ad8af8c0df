"""CPU oracle: restatement of the reference's social-force stepping path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Plain numpy/scipy
float64, struct-of-arrays over agents; the per-agent parts are scalar Python
(like the reference), the all-pairs part is an N x N numpy evaluation (the
reference builds N^2 x N^2 temporaries by accident, intersection.py:711-720).

Every function cites the reference lines it follows; paths are relative to
``/root/reference/src/cyclistsocialforce/``.

Pinning status (DESIGN.md section "Oracle"):
  * TwoDBicycle, PlanarPointBicycle, Bicycle, pair force, FOV mask, nav
    machine, road force: pinned against the reference's own code run through
    ``oracle/ref_harness.py`` (tests/test_oracle_vs_reference.py, and the
    committed vectors under tests/golden/).
  * InvPendulumBicycle and BalancingRiderBicycle: the reference's own step
    logic is pinned the same way, but the third-party arithmetic underneath
    (python-control ``forced_response``/``place``, ``bicycleparameters``
    matrices) is absent from this image and was restated from the published
    algorithms -> **parity unpinned** with respect to those packages.
"""
from __future__ import annotations

import copy
import math
from types import SimpleNamespace

import numpy as np

TWO_PI = 2.0 * np.pi

# ---------------------------------------------------------------------------
# parameter defaults (parameters.py:430-451, :780-802, :1180-1186, :1216-1228,
# :1429-1472, :368-375)
# ---------------------------------------------------------------------------
_VEHICLE = dict(
    t_s=0.01, d_arrived_inter=2.0, d_arrived_stop=2.0, v_max_stop=0.1,
    v_max_harddecel=2.5, hfov=2 * np.pi,
    f_0=7.0, e_0=0.995, e_1=0.7, sigma_0=0.5, sigma_1=5.0, sigma_2=0.3, sigma_3=4.9,
)
_BICYCLE = dict(
    _VEHICLE, v_max_riding=(-1.0, 10.0), v_desired_default=5.0, p_decay=5.0, p_0=30.0,
    hfov=np.pi * 2 / 3, v_max_stop=0.6, l=1.0, l_1=0.5, l_2=0.5, delta_max=1.4,
    a_max=(-10.0, 10.0), a_desired_default=(-5.0, 5.0), k_p_v=10.0, k_p_delta=10.0, g=9.81,
)
_INVPEND = dict(
    _BICYCLE, v_max_riding=(-1.0, 7.0), a_max=(-3.0, 1.0), a_desired_default=(-1.0, 0.5),
    h=1.0, m=87.0, i_bike_longlong=3.28, i_steer_vertvert=0.07, c_steer=50.0,
    v_max_walk=1.5, delta_max_walk=0.174,
)
_PLANARPOINT = dict(_BICYCLE, k_psi=2.0)

#: balanceassistv1_with_averagerider, data/bicycleparams/balanceassist_bikeparams.py:11-39
BALANCEASSIST = dict(
    IBxx=16.136560964517308, IBxz=-2.5375819134691833, IByy=18.98228436804581,
    IBzz=4.308368614306412, IFxx=0.0995, IFyy=0.1902, IHxx=0.2984, IHxz=-0.038,
    IHyy=0.257, IHzz=0.0566, IRxx=0.1023, IRyy=0.1887, c=0.042, g=9.81, lam=0.255,
    mB=91.50000000000003, mF=2.235, mH=4.3, mR=4.085, rF=0.35231, rR=0.34895, v=1.0,
    w=1.113, xB=0.373106714751133, xH=0.921, yB=0.0, zB=-0.9697039390081493, zH=-0.86,
)

#: linear pole(v) regressions of the packaged rider-behaviour models, produced
#: by the reference's PoleModel.get_component_mean_function
#: (controlbehavior.py:1601-1650) via tests/golden/make_golden.py.
#: features = [p0_real, p1_real, p1_imag, p2_real, p2_imag]
POLE_REGRESSIONS = {
    ("BR1", 0): ([-1.8535775147013691, -0.17961038648194527, 1.1730293099635356,
                  0.6329343335691049, 2.3198026534882037],
                 [-1.6314600197495845, -0.15971107558265726, 0.14249117463843872,
                  -0.44818335654213604, 0.9166444997604464]),
    ("BR1", 1): ([-0.35234247428779364, -0.10273325648260068, 1.738236648818369,
                  -0.47097770394470295, 7.811765217705167],
                 [-0.588378287364706, -0.28745435083339654, 0.3111051504915779,
                  -0.7008198641061155, 0.06086016583769994]),
    ("BR0", 0): ([7.477367764370239, -0.6066675056522426, 1.7881981548329726,
                  -1.32823781934098, 5.327111219864689],
                 [-7.589580229524327, -0.10886032204606368, 0.04106272039749894,
                  -0.02666449568231668, 0.08910709292445088]),
}
_BALANCINGRIDER = dict(
    _BICYCLE, l=BALANCEASSIST["w"], l_1=BALANCEASSIST["w"] / 2, l_2=BALANCEASSIST["w"] / 2,
    pole_model=("BR1", 0), bike=BALANCEASSIST,
)

MODELS = ("twod", "invpendulum", "balancingrider", "planarpoint", "bicycle", "uncontrolled")
N_STATES = dict(twod=5, invpendulum=6, balancingrider=8, planarpoint=4, bicycle=5,
                uncontrolled=4)
TRAJ_LEN = 3000          # int(30 / t_s), vehicle.py:159
HIST = 100               # int(1 / t_s), vehicle.py:1487


def default_params(model: str, **over) -> SimpleNamespace:
    base = dict(twod=_INVPEND, invpendulum=_INVPEND, balancingrider=_BALANCINGRIDER,
                planarpoint=_PLANARPOINT, bicycle=_BICYCLE, uncontrolled=_VEHICLE)[model]
    d = copy.deepcopy(dict(base))
    d.update(over)
    p = SimpleNamespace(**d)
    if model in ("twod", "invpendulum"):
        p.l = p.l_1 + p.l_2
        # parameters.py tau_1_squared: (I_ll + m h^2) / (m g h)
        p.tau_1_squared = (p.i_bike_longlong + p.m * p.h**2) / (p.m * p.g * p.h)
    return p


# ---------------------------------------------------------------------------
# utils.py
# ---------------------------------------------------------------------------
def limit_angle(theta):
    """utils.py:124-139 -- wrap to (-pi, pi]."""
    if isinstance(theta, np.ndarray):
        theta = np.floor(theta / TWO_PI) * (-TWO_PI) + theta
        theta = np.where(theta > np.pi, theta - TWO_PI, theta)
        theta = np.where(theta < -np.pi, theta + TWO_PI, theta)
        return theta
    theta = math.floor(theta / TWO_PI) * (-TWO_PI) + theta
    if theta > np.pi:
        theta = theta - TWO_PI
    elif theta < -np.pi:
        theta = theta + TWO_PI
    return theta


def angle_difference(a1, a2):
    """utils.py:151-182 -- signed shortest rotation a1 -> a2 (ties -> +)."""
    if isinstance(a1, np.ndarray) or isinstance(a2, np.ndarray):
        a1, a2 = np.broadcast_arrays(np.asarray(a1, float), np.asarray(a2, float))
        da = np.where(a1 > a2, a1 - a2, a2 - a1)
        da = np.where(da > np.pi, TWO_PI - da, da)
        t1 = np.abs(limit_angle(a1 - da) - a2)
        t2 = np.abs(limit_angle(a1 + da) - a2)
        return np.where(t1 < t2, -da, da)
    da = a1 - a2 if a1 > a2 else a2 - a1
    if da > np.pi:
        da = TWO_PI - da
    t1 = abs(limit_angle(a1 - da) - a2)
    t2 = abs(limit_angle(a1 + da) - a2)
    return -da if t1 < t2 else da


def cart2polar(x, y):
    """utils.py:185-194."""
    rho = np.sqrt(np.power(x, 2) + np.power(y, 2))
    with np.errstate(invalid="ignore", divide="ignore"):
        phi = np.arccos(x / rho)
    phi = np.where(y < 0, -phi, phi)
    return rho, phi


def thresh(x, minmax):
    """utils.py:204-227."""
    return np.maximum(np.minimum(x, minmax[1]), minmax[0])


def limit_magnitude(x, y, r):
    """utils.py:56-86 (returns copies)."""
    x = np.array(x, float)
    y = np.array(y, float)
    rin = np.sqrt(x**2 + y**2)
    ids = rin > r
    if np.any(rin):
        x[ids] = x[ids] * r[ids] / rin[ids]
        y[ids] = y[ids] * r[ids] / rin[ids]
    return x, y


# ---------------------------------------------------------------------------
# all-pairs repulsive force  (a2, a3, a4)
# ---------------------------------------------------------------------------
def field_params_array(plist):
    """(S, 8) array [f_0, e_0, e_1, sigma_0..3, hfov] for a list of params."""
    return np.array([[p.f_0, p.e_0, p.e_1, p.sigma_0, p.sigma_1, p.sigma_2, p.sigma_3, p.hfov]
                     for p in plist], float)


def twod_field(x0, y0, psi0, fp, x, y, psi, literal=False):
    """TwoDBicycle.calcRepulsiveForce, vehicle.py:1584-1648.

    Sources (x0,y0,psi0) broadcast against targets (x,y,psi); ``fp`` = columns of
    ``field_params_array`` broadcastable the same way.  Returns Fx, Fy.

    ``literal=True`` reproduces the reference's arithmetic operation by operation,
    including its far-field failure: for rho*q/sigma > ~372 the squares in
    ``F = sqrt(Fx**2 + Fy**2)`` (:1644) underflow to 0 and ``P*Fx/F`` is 0/0 = NaN,
    i.e. any crowd wider than ~100 m yields NaN forces in the reference.  The default
    evaluates the same expression with the common factor P taken out of the
    normalisation (identical to 1 ulp wherever the reference is finite, and the
    correct limit -- P times a unit vector -- where it is not).
    """
    f0, e0, e1, s0, s1, s2_, s3 = (fp[..., k] for k in range(7))
    psi_rel = psi0 - psi
    sin2 = np.sin(psi_rel) ** 2
    vd0 = s0 + s1 * sin2
    vd1 = s2_ + s3 * sin2
    e = e0 - e1 * sin2
    dx = x - x0
    dy = y - y0
    rho, phi1 = cart2polar(dx, dy)
    phi = limit_angle(np.asarray(phi1 - psi0))
    cosphi = np.cos(phi)
    sinphi = np.sin(phi)
    with np.errstate(invalid="ignore", divide="ignore", under="ignore"):
        sigma = vd0 - vd1 * np.sqrt((1 - cosphi) / 2)
        dsigm = -vd1 * np.sqrt((1 + cosphi) / 2) * np.sign(phi) / 2
        q = np.sqrt(1 - (e * cosphi) ** 2)
        P = f0 * np.exp(-rho * q / sigma)
        one = P if literal else 1.0
        Frho = one * q / sigma
        Fphi = -one * ((1 - (e * cosphi) ** 2) * dsigm - e**2 * sinphi * cosphi * sigma) / (
            sigma**2 * q)
        Fx = Frho * np.cos(phi1) - Fphi * np.sin(phi1)
        Fy = Frho * np.sin(phi1) + Fphi * np.cos(phi1)
        F = np.sqrt(Fx**2 + Fy**2)
        Fx = P * Fx / F
        Fy = P * Fy / F
    zero = np.broadcast_to(f0 == 0.0, Fx.shape)          # vehicle.py:1592-1593
    return np.where(zero, 0.0, Fx), np.where(zero, 0.0, Fy)


def bicycle_field(x0, y0, psi0, v0, p, x, y):
    """Bicycle.calcRepulsiveForce (v0.1 elliptic potential), vehicle.py:1054-1147."""
    with np.errstate(invalid="ignore", divide="ignore"):
        e = np.minimum(np.power(v0 / p.v_max_riding[1], 0.1), 0.7)
        dx = x - x0
        dy = y - y0
        rho, phi = cart2polar(dx, dy)
        phi0 = phi - psi0
        b = (1 / (np.sqrt(1 - e**2) * p.p_decay)) * rho * (1 - e * np.cos(phi0))
        P = p.p_0 * np.exp(-b) / p.p_decay
        Frho0 = P * ((1 - e * np.cos(phi0)) / np.sqrt(1 - e**2))
        Fphi0 = P * ((e * np.sin(phi0)) / np.sqrt(1 - e**2))
        Fx = Frho0 * np.cos(phi) - Fphi0 * np.sin(phi)
        Fy = Frho0 * np.sin(phi) + Fphi0 * np.cos(phi)
    return Fx, Fy


def tracked_mask(x, y, psi, hfov_src, p2r=False, tgt=None, return_margin=False):
    """``not get_untracked_foes()``, intersection.py:690-745, as [source i, target j].

    tracked[i, j] = i != j  and  |angDiff(psi_j, atan2(y_i-y_j, x_i-x_j))| <= hfov_i/2
                    [and rel-azimuth <= 0 under "p2r"].
    ``tgt`` restricts targets to an index array (columns).  ``margin`` is
    hfov_i/2 - |rel| (distance to the FOV boundary, rad).
    """
    x, y, psi = (np.asarray(a, float) for a in (x, y, psi))
    n = x.shape[0]
    tj = np.arange(n) if tgt is None else np.asarray(tgt)
    dxs = x[:, None] - x[None, tj]
    dys = y[:, None] - y[None, tj]
    az = limit_angle(np.arctan2(dys, dxs))
    rel = angle_difference(np.broadcast_to(psi[None, tj], az.shape).copy(), az)
    hf = np.broadcast_to(np.asarray(hfov_src, float), (n,))[:, None] / 2
    untracked = np.abs(rel) > hf
    untracked |= np.arange(n)[:, None] == tj[None, :]
    if p2r:
        untracked |= rel > 0
    if return_margin:
        return ~untracked, hf - np.abs(rel), rel
    return ~untracked


def pair_forces(x, y, psi, fparams, p2r=False, tgt=None, chunk=1024, src_v=None,
                src_kind=None, bicycle_params=None, return_margin=False):
    """Sum over sources of the masked pair force: intersection.py:788-823, :841-843.

    ``fparams`` is (8,) (homogeneous) or (N, 8) per *source* (the reference takes
    field parameters and hfov from the source vehicle, intersection.py:733-735,
    vehicle.py:1587-1612).  ``src_kind[i] == 1`` selects the v0.1 ``Bicycle``
    field for source i (needs ``src_v`` and ``bicycle_params``).
    Returns Frep (T, 2) for the requested targets (all by default).
    With N == 1 the reference skips the pair stage entirely (:813) -> zeros.
    """
    x, y, psi = (np.asarray(a, float) for a in (x, y, psi))
    n = x.shape[0]
    tj = np.arange(n) if tgt is None else np.asarray(tgt)
    out = np.zeros((tj.shape[0], 2))
    min_margin = np.full(tj.shape[0], np.inf)
    if n <= 1:
        return (out, min_margin) if return_margin else out
    fp = np.broadcast_to(np.asarray(fparams, float), (n, 8))
    for a in range(0, tj.shape[0], chunk):
        t = tj[a:a + chunk]
        tr, margin, _ = tracked_mask(x, y, psi, fp[:, 7], p2r=p2r, tgt=t, return_margin=True)
        Fx, Fy = twod_field(x[:, None], y[:, None], psi[:, None], fp[:, None, :],
                            x[None, t], y[None, t], psi[None, t])
        if src_kind is not None and np.any(np.asarray(src_kind) == 1):
            k1 = np.asarray(src_kind) == 1
            bx, by = bicycle_field(x[k1, None], y[k1, None], psi[k1, None],
                                   np.asarray(src_v, float)[k1, None], bicycle_params,
                                   x[None, t], y[None, t])
            Fx[k1], Fy[k1] = bx, by
        Fx = np.where(tr, Fx, 0.0)
        Fy = np.where(tr, Fy, 0.0)
        out[a:a + chunk, 0] = Fx.sum(axis=0)
        out[a:a + chunk, 1] = Fy.sum(axis=0)
        notself = np.arange(n)[:, None] != t[None, :]
        min_margin[a:a + chunk] = np.where(notself, np.abs(margin), np.inf).min(axis=0)
    return (out, min_margin) if return_margin else out


def road_forces(x, y, vertices, F_0=0.05, sigma=3.0):
    """RoadEdge.calcRepulsiveForce, intersection.py:226-242 (one edge)."""
    x = np.asarray(x, float).ravel()
    y = np.asarray(y, float).ravel()
    v = np.asarray(vertices, float)
    r = np.sqrt((v[:, 0][None, :] - x[:, None]) ** 2 + (v[:, 1][None, :] - y[:, None]) ** 2)
    erx = (v[:, 0][None, :] - x[:, None]) / r
    ery = (v[:, 1][None, :] - y[:, None]) / r
    F = -F_0 * r ** (-sigma)
    return np.sum(F * erx, axis=1), np.sum(F * ery, axis=1)


# ---------------------------------------------------------------------------
# interpolating parametric cubic B-spline  (scipy FITPACK, as the reference)
# ---------------------------------------------------------------------------
def spline_samples(px, py, n=20):
    """splprep(s=0) + 3 x splev at linspace(0,1,n), vehicle.py:1496-1512."""
    from scipy import interpolate
    tck, _ = interpolate.splprep((px, py), s=0.0)
    u = np.linspace(0, 1, n)
    xs, ys = interpolate.splev(u, tck)
    dxs, dys = interpolate.splev(u, tck, der=1)
    d2xs, d2ys = interpolate.splev(u, tck, der=2)
    return np.c_[xs, ys, dxs, dys, d2xs, d2ys]


# ---------------------------------------------------------------------------
# Whipple-Carvallo + pole placement (BalancingRider)
# ---------------------------------------------------------------------------
def meijaard_canonical(p: dict):
    """M, C1, K0, K2 of  M q'' + v C1 q' + (g K0 + v^2 K2) q = f, q=[phi,delta].

    Restated from Meijaard, Papadopoulos, Ruina, Schwab (2007), Appendix A
    (SURVEY.md Appendix A.5).
    """
    w, c, lam = p["w"], p["c"], p["lam"]
    rR, mR, IRxx, IRyy = p["rR"], p["mR"], p["IRxx"], p["IRyy"]
    xB, zB, mB = p["xB"], p["zB"], p["mB"]
    IBxx, IBxz, IBzz = p["IBxx"], p["IBxz"], p["IBzz"]
    xH, zH, mH = p["xH"], p["zH"], p["mH"]
    IHxx, IHxz, IHzz = p["IHxx"], p["IHxz"], p["IHzz"]
    rF, mF, IFxx, IFyy = p["rF"], p["mF"], p["IFxx"], p["IFyy"]
    sl, cl = np.sin(lam), np.cos(lam)

    mT = mR + mB + mH + mF
    xT = (xB * mB + xH * mH + w * mF) / mT
    zT = (-rR * mR + zB * mB + zH * mH - rF * mF) / mT
    ITxx = IRxx + IBxx + IHxx + IFxx + mR * rR**2 + mB * zB**2 + mH * zH**2 + mF * rF**2
    ITxz = IBxz + IHxz - mB * xB * zB - mH * xH * zH + mF * w * rF
    ITzz = IRxx + IBzz + IHzz + IFxx + mB * xB**2 + mH * xH**2 + mF * w**2
    mA = mH + mF
    xA = (xH * mH + w * mF) / mA
    zA = (zH * mH - rF * mF) / mA
    IAxx = IHxx + IFxx + mH * (zH - zA) ** 2 + mF * (rF + zA) ** 2
    IAxz = IHxz - mH * (xH - xA) * (zH - zA) + mF * (w - xA) * (rF + zA)
    IAzz = IHzz + IFxx + mH * (xH - xA) ** 2 + mF * (w - xA) ** 2
    uA = (xA - w - c) * cl - zA * sl
    IAll = mA * uA**2 + IAxx * sl**2 + 2 * IAxz * sl * cl + IAzz * cl**2
    IAlx = -mA * uA * zA + IAxx * sl + IAxz * cl
    IAlz = mA * uA * xA + IAxz * sl + IAzz * cl
    mu = c / w * cl
    SR, SF = IRyy / rR, IFyy / rF
    ST = SR + SF
    SA = mA * uA + mu * mT * xT
    M = np.array([[ITxx, IAlx + mu * ITxz],
                  [IAlx + mu * ITxz, IAll + 2 * mu * IAlz + mu**2 * ITzz]])
    K0 = np.array([[mT * zT, -SA], [-SA, -SA * sl]])
    K2 = np.array([[0.0, (ST - mT * zT) / w * cl],
                   [0.0, (SA + SF * sl) / w * cl]])
    C1 = np.array([[0.0, mu * ST + SF * cl + ITxz / w * cl - mu * mT * zT],
                   [-(mu * ST + SF * cl), IAlz / w * cl + mu * (SA + ITzz / w * cl)]])
    return M, C1, K0, K2


def balancingrider_matrices(bike: dict, v: float):
    """A (5x5), B (5,) of x=[phi,delta,phidot,deltadot,psi], steer-torque input.
    dynamics.py:511-538 (+ :296-302)."""
    M, C1, K0, K2 = meijaard_canonical(bike)
    Minv = np.linalg.inv(M)
    A = np.zeros((5, 5))
    A[0:2, 2:4] = np.eye(2)
    A[2:4, 0:2] = -Minv @ (bike["g"] * K0 + v**2 * K2)
    A[2:4, 2:4] = -Minv @ (v * C1)
    coslam = np.cos(bike["lam"])
    A[4, 1] = coslam / bike["w"] * v
    A[4, 3] = coslam * bike["c"] / bike["w"]
    B = np.zeros(5)
    B[2:4] = Minv[:, 1]
    return A, B


def balancingrider_poles(pole_model, v):
    """BalancingRiderBicycleParameters.update_control_params, parameters.py:1403-1411."""
    icpt, coef = POLE_REGRESSIONS[tuple(pole_model)]
    f = np.asarray(icpt) + np.asarray(coef) * v
    poles = [f[0] + 0j]
    i = 1
    while i < len(f):
        poles.append(f[i] + 1j * f[i + 1])
        poles.append(f[i] - 1j * f[i + 1])
        i += 2
    return np.array(poles)


def ps_poles(f):
    """[p0_real, p1_real, p1_imag, p2_real, p2_imag] -> the five closed-loop poles (controlbehavior.py:65-89)."""
    return np.array([f[0] + 0j, f[1] + 1j * f[2], f[1] - 1j * f[2], f[3] + 1j * f[4], f[3] - 1j * f[4]])


def place_gain(A, B, poles):
    """ct.place == scipy.signal.place_poles(method='YT').gain_matrix (dynamics.py:1209)."""
    import scipy.signal
    return scipy.signal.place_poles(A, B.reshape(-1, 1), poles, method="YT").gain_matrix[0]


# ---------------------------------------------------------------------------
# agent groups
# ---------------------------------------------------------------------------
class Agents:
    """One model type, struct of arrays (all float64)."""

    def __init__(self, model, s0, params=None, v_desired=None, destqueue=None,
                 uncontrolled_traj=None):
        assert model in MODELS
        self.model = model
        self.p = params if params is not None else default_params(model)
        ns = N_STATES[model]
        s0 = np.atleast_2d(np.asarray(s0, float))[:, :ns].copy()
        self.n = s0.shape[0]
        self.s = s0
        self.s[:, 2] = [limit_angle(float(a)) for a in s0[:, 2]]      # vehicle.py:155
        self.i = np.zeros(self.n, int)
        self.vd_default = np.full(self.n, getattr(self.p, "v_desired_default", 0.0), float)
        if v_desired is not None:
            self.vd_default[:] = v_desired
        # destination queue: first entry = start position (vehicle.py:183-185)
        self.queues = [np.array([[s0[k, 0], s0[k, 1], 0.0]]) for k in range(self.n)]
        self.ptr = np.zeros(self.n, int)
        self.znav = np.zeros((self.n, 3), bool)
        self.znav[:, 0] = True
        self.znavparams = np.zeros((self.n, 4))
        # ring of past states (vehicle.py:159-160)
        self.traj = np.zeros((self.n, ns, TRAJ_LEN))
        self.traj[:, :, 0] = self.s
        self.force = np.zeros((self.n, 2))
        self.fdest = np.zeros((self.n, 2))
        self.frep = np.zeros((self.n, 2))
        if destqueue is not None:
            for k in range(self.n):
                q = np.asarray(destqueue[k], float)
                self.set_destinations(k, q[:, 0], q[:, 1], q[:, 2] if q.shape[1] > 2 else None)
        if model == "invpendulum":
            # vehicle.py:1728-1736
            self.x = np.stack([self.s[:, 4], 0 * self.s[:, 4], self.s[:, 5], 0 * self.s[:, 4],
                               self.s[:, 2]], axis=1)
            self.zrid = np.zeros((self.n, 2), bool)
            walk = s0[:, 3] < self.p.v_max_walk
            self.zrid[walk, 1] = True
            self.zrid[~walk, 0] = True
        if model == "balancingrider":
            # dynamics.py:361-399 (csf -> bike frame), :305-306
            s = self.s
            self.x = np.stack([s[:, 5], -s[:, 4], s[:, 7], -s[:, 6], -s[:, 2], s[:, 0], -s[:, 1]],
                              axis=1)
            self.v = s[:, 3].copy()
            # stochastic rider behaviour (parameters.py:1311, :1398-1402): per-rider pole features, the
            # speed they were drawn at and the number of draws consumed from the rider's random stream
            self.br_feats = np.zeros((self.n, 5))
            self.br_vlast = np.full(self.n, -10000.0)
            self.br_draws = np.zeros(self.n, int)
            self.gains = np.stack([self._br_gains(float(v), k) for k, v in enumerate(self.v)])
        if model == "planarpoint":
            self.x = np.stack([self.s[:, 2], self.s[:, 0], self.s[:, 1]], axis=1)   # dynamics.py:987-991
            self.v = self.s[:, 3].copy()
        if model == "uncontrolled":
            self.utraj = uncontrolled_traj          # list of (4, T) arrays or None

    # -- destinations -------------------------------------------------------
    def set_destinations(self, k, x, y, stop=None, reset=False):
        """Vehicle.setDestinations, vehicle.py:606-647."""
        x = np.array([x], float).flatten()
        y = np.array([y], float).flatten()
        stop = np.zeros_like(x) if stop is None else np.array([stop], float).flatten()
        if reset:
            self.queues[k] = np.c_[x, y, stop]
            self.ptr[k] = 0
        else:
            self.queues[k] = np.vstack((self.queues[k], np.c_[x, y, stop]))

    def dest(self, k):
        return self.queues[k][self.ptr[k]]

    def is_last_dest(self, k):
        return self.ptr[k] + 1 >= self.queues[k].shape[0]          # vehicle.py:537-543

    def dest_distance(self, k):
        d = self.dest(k)                                             # vehicle.py:596-604
        return math.sqrt((d[0] - self.s[k, 0]) ** 2 + (d[1] - self.s[k, 1]) ** 2)

    def update_destination(self, k):
        """Vehicle.updateDestination, vehicle.py:545-594."""
        q = self.queues[k]
        dnext = self.dest_distance(k)
        if self.znav[k, 1] or self.znav[k, 2]:
            return
        if dnext <= self.p.d_arrived_inter:
            self.ptr[k] = min(self.ptr[k] + 1, q.shape[0] - 1)
        if self.ptr[k] < q.shape[0] - 1:
            nn = q[self.ptr[k] + 1]
            dnn = math.sqrt((nn[0] - self.s[k, 0]) ** 2 + (nn[1] - self.s[k, 1]) ** 2)
            if dnn < dnext:
                self.ptr[k] += 1

    def update_nav_state(self, k, stop):
        """Vehicle.updateNavState, vehicle.py:354-457 -> (v_d, d_dest)."""
        p = self.p
        kk = 1.5
        z = self.znav[k].copy()
        v = self.s[k, 3]
        if z[0]:
            d0 = 0.5 * (p.v_max_harddecel**2 - v**2) / p.a_desired_default[0]
            d1 = 0.5 * -p.v_max_harddecel**2 / p.a_max[0]
        else:
            d0, d1 = self.znavparams[k, 1], self.znavparams[k, 2]
        ddest = self.dest_distance(k)
        x0 = bool(stop)
        x1 = ddest <= kk * (d0 + d1)
        x2 = ddest <= p.d_arrived_stop
        x3 = v <= p.v_max_stop
        n0 = (not x0) or (x0 and (not x1) and ((z[0] and not x2) or z[1]))
        n1 = x0 and ((z[0] and (((not x2) and x1) or (x2 and not x3)))
                     or (z[1] and x1 and ((not x2) or (not x3))))
        n2 = x0 and (((z[0] or z[1]) and x2 and x3) or z[2])
        self.znav[k] = (n0, n1, n2)
        if z[0] and n1:
            self.znavparams[k] = (v, d0, d1, self.i[k])
        zp = self.znavparams[k]
        if n0:
            vd = self.vd_default[k]
        elif n1:
            if ddest < kk * zp[2]:
                vd = p.v_max_harddecel / zp[2] * ddest * 1 / kk
            else:
                vd = (zp[0] - p.v_max_harddecel) / zp[1] * (ddest - zp[2]) * 1 / kk \
                    + p.v_max_harddecel
        elif n2:
            vd = 0.0
        else:
            raise RuntimeError("Invalid navigation state")
        return vd, ddest

    # -- destination forces -------------------------------------------------
    def dest_force_direct_field(self, k):
        """Bicycle.calcDestinationForceField / calc_direct_approach_dest_force,
        vehicle.py:1168-1187, :2096-2108."""
        self.update_destination(k)
        d = self.dest(k)
        vd, ddest = self.update_nav_state(k, d[2])
        if ddest > 0:
            return (-vd * (self.s[k, 0] - d[0]) / ddest, -vd * (self.s[k, 1] - d[1]) / ddest)
        return 0.0, 0.0

    def dest_force_twod(self, k):
        """TwoDBicycle.calcDestinationForce, vehicle.py:1443-1558."""
        self.update_destination(k)
        d = self.dest(k)
        vd, ddest = self.update_nav_state(k, d[2])
        s = self.s[k]
        i = self.i[k]
        if i == 0:
            return vd * math.cos(s[2]), vd * math.sin(s[2])
        if self.znav[k, 2]:
            return 0.0, 0.0
        q = self.queues[k]
        tr = self.traj[k]
        last = self.is_last_dest(k)
        if not last:
            idest = np.arange(self.ptr[k], min(self.ptr[k] + 4, q.shape[0]), dtype=int)
            px = np.r_[tr[0, (i - 1, i)], q[idest, 0]]
            py = np.r_[tr[1, (i - 1, i)], q[idest, 1]]
        else:
            ispl = (max(0, i - HIST), i - 1, i)
            px = np.r_[tr[0, ispl], d[0]]
            py = np.r_[tr[1, ispl], d[1]]
        S = spline_samples(px, py, 20)
        if last:
            i_s = int(np.argmin((S[:, 0] - s[0]) ** 2 + (S[:, 1] - s[1]) ** 2))
        else:
            i_s = 1
        i_p = i_s + (5 if d[2] else 3)
        if i_p < 20:
            with np.errstate(divide="ignore"):
                R = math.sqrt(S[i_s, 2] ** 2 + S[i_s, 3] ** 2) ** 3 / np.abs(
                    S[i_s, 2] * S[i_s, 5] - S[i_s, 3] * S[i_s, 4])
            v = max(2.5, math.sqrt(10 * (2 * np.pi / 360) * self.p.g * R))
            v = min(v, vd)
            ddx = S[i_p, 0] - S[i_s, 0]
            ddy = S[i_p, 1] - S[i_s, 1]
            temp = v / math.sqrt(ddx**2 + ddy**2)
            return temp * ddx, temp * ddy
        return self.dest_force_direct_field(k)           # vehicle.py:1556 (super())

    def calc_destination_force(self, k):
        """Per-class dispatch."""
        m = self.model
        if m in ("twod", "invpendulum"):
            return self.dest_force_twod(k)
        if m == "bicycle":
            return self.dest_force_direct_field(k)        # vehicle.py:1189-1194
        if m == "balancingrider":
            self.update_destination(k)                    # vehicle.py:295-297
            return self.dest_force_direct_field(k)
        if m == "planarpoint":
            self.update_destination(k)                    # vehicle.py:295-297
            return self.dest_force_twod(k)                # vehicle.py:2025
        return 0.0, 0.0                                   # uncontrolled, vehicle.py:986-987

    # -- control + dynamics -------------------------------------------------
    def _control(self, k, Fx, Fy):
        """Bicycle.control + PIDcontroller (kp only), vehicle.py:1218-1245."""
        p = self.p
        s = self.s[k]
        d = self.dest(k)
        theta = math.atan2(Fy, Fx)
        v = math.sqrt(Fx**2 + Fy**2)
        ddest = math.sqrt((d[0] - s[0]) ** 2 + (d[1] - s[1]) ** 2)
        if ddest < 3 and self.is_last_dest(k):
            v = (v / 3) * ddest
        target = angle_difference(s[2], theta)
        ddelta = angle_difference(s[4], target)
        dv = v - s[3]
        return p.k_p_v * dv, p.k_p_delta * ddelta

    def _move(self, k, a, ddelta):
        """Bicycle.move, vehicle.py:1247-1272."""
        p = self.p
        s = self.s[k]
        a = thresh(a, p.a_max)
        delta = limit_angle(s[4] + p.t_s * ddelta)
        v = s[3] + p.t_s * a
        delta = thresh(delta, (-p.delta_max, p.delta_max))
        v = thresh(v, p.v_max_riding)
        theta = limit_angle(s[2] + p.t_s * v * math.tan(delta) / p.l)
        s[1] = s[1] + p.t_s * v * math.sin(theta)
        s[0] = s[0] + p.t_s * v * math.cos(theta)
        s[2], s[3], s[4] = theta, v, delta

    def _record(self, k, wrap=True):
        self.i[k] += 1
        if wrap:
            self.i[k] %= TRAJ_LEN
        self.traj[k, :, self.i[k] % TRAJ_LEN] = self.s[k]

    def step_bicycle(self, k, Fx, Fy):
        """Bicycle.step, vehicle.py:1274-1289."""
        a, od = self._control(k, Fx, Fy)
        self._move(k, a, od)
        self._record(k)

    def step_twod(self, k, Fx, Fy):
        """TwoDBicycle.step, vehicle.py:1386-1414."""
        if self.znav[k, 2]:
            self.s[k, 3:6] = 0
        else:
            a, od = self._control(k, Fx, Fy)
            self._move(k, a, od)
        self._record(k)

    def invpend_closed_loop(self, v):
        """A_c, B_c: vehicle.py:1738-1786, parameters.py:1832-1892."""
        p = self.p
        K_tau_2 = (v * p.l_2) / (p.g * p.l)
        K = v**2 / (p.g * p.l)
        tau_3 = p.l / v
        A = np.zeros((5, 5))
        A[0, 1] = 1
        A[1, 1] = -p.c_steer / p.i_steer_vertvert
        A[2, 3] = 1
        A[3, 0] = -K / p.tau_1_squared
        A[3, 1] = -K_tau_2 / p.tau_1_squared
        A[3, 2] = 1 / p.tau_1_squared
        A[4, 0] = 1 / tau_3
        B = np.array([0, 1 / p.i_steer_vertvert, 0, 0, 0])
        kx, ku = invpend_gains(v)
        return A - B[:, None] @ kx[None, :], ku * B

    def step_invpendulum(self, k, Fx, Fy):
        """InvPendulumBicycle.step, vehicle.py:1883-1950."""
        import scipy.linalg
        p = self.p
        s = self.s[k]
        i = self.i[k]
        # updateRidingState, :1932-1950
        cvwalk = s[3] < p.v_max_walk
        imin = max(0, int(i - 1 / p.t_s))
        win = self.traj[k, 4, imin:i + 1]
        cdelta = bool(np.all(-p.delta_max_walk < win) and np.all(p.delta_max_walk > win))
        z0 = (not cvwalk) and ((self.zrid[k, 1] and cdelta) or self.zrid[k, 0])
        self.zrid[k] = (z0, not z0)
        if self.znav[k, 2]:
            s[3:6] = 0
        elif self.zrid[k, 0]:
            # step_pos (:1850-1881) first, with the old psi
            vd = math.sqrt(Fx**2 + Fy**2)
            a = thresh(p.k_p_v * (vd - s[3]), p.a_max)
            v = thresh(s[3] + p.t_s * a, p.v_max_riding)
            ynew = s[1] + p.t_s * v * math.sin(s[2])
            xnew = s[0] + p.t_s * v * math.cos(s[2])
            s[0], s[1], s[3] = xnew, ynew, v
            # step_yaw (:1810-1848) with the updated speed
            Ac, Bc = self.invpend_closed_loop(s[3])
            psi_d = math.atan2(Fy, Fx)
            n = 5
            M = np.zeros((n + 2, n + 2))
            M[:n, :n] = Ac * p.t_s
            M[:n, n] = Bc * p.t_s
            M[n, n + 1] = 1.0
            eM = scipy.linalg.expm(M)
            Ad, Bd1 = eM[:n, :n], eM[:n, n + 1]
            Bd0 = eM[:n, n] - Bd1
            xn = Ad @ self.x[k] + Bd0 * psi_d + Bd1 * psi_d
            self.x[k] = xn
            s[2], s[4], s[5] = limit_angle(xn[4]), limit_angle(xn[0]), limit_angle(xn[2])
        else:
            s[3] = p.v_max_walk
            s[5] = 0
            a, od = self._control(k, Fx, Fy)
            self._move(k, a, od)
            self.x[k] = (s[4], 0, s[5], 0, s[2])
        self._record(k)

    def _br_gains(self, v, k=0):
        """BalancingRiderDynamics._get_gains, dynamics.py:602-615 (update_control_params(v), then place)."""
        if getattr(self.p, "fixed_gains", None) is not None:       # dynamics.py:606-607
            return np.asarray(self.p.fixed_gains, float)
        A, B = balancingrider_matrices(self.p.bike, v)
        if getattr(self.p, "fixed_poles", None) is not None:       # parameters.py:1311-1314: controlparam_fix
            return place_gain(A, B, np.asarray(self.p.fixed_poles, complex))
        if getattr(self.p, "stochastic", False):
            # parameters.py:1398-1402: new, independent poles when the speed has moved by more than the
            # threshold since the last draw; otherwise the rider keeps its poles
            if abs(v - self.br_vlast[k]) > self.p.resample_thresh:
                from . import pole_sampling as ps
                m = ps.load_model(self.p.pole_model_file)
                f, used = ps.sample_features(m, v, self.p.seed, self.p.agent_offset + k, int(self.br_draws[k]))
                self.br_feats[k] = f
                self.br_draws[k] += used
                self.br_vlast[k] = v
            return place_gain(A, B, ps_poles(self.br_feats[k]))
        return place_gain(A, B, balancingrider_poles(self.p.pole_model, v))

    def step_balancingrider(self, k, Fx, Fy):
        """BalancingRiderDynamics.step, dynamics.py:674-705 (closed-form midpoint;
        the reference solves the same linear-in-x_br system with MINPACK lm)."""
        p = self.p
        vd = math.sqrt(Fx**2 + Fy**2)
        a = thresh(p.k_p_v * (vd - self.v[k]), p.a_max)
        v = thresh(self.v[k] + p.t_s * a, p.v_max_riding)
        vbar = (v + self.s[k, 3]) / 2
        if v != self.v[k]:
            self.gains[k] = self._br_gains(vbar, k)
        g = self.gains[k]
        x = self.x[k]
        psi_F = limit_angle(math.atan2(-Fy, Fx))                      # :661-671
        psi_c = x[4] + angle_difference(x[4], psi_F)
        A, B = balancingrider_matrices(p.bike, vbar)
        Ac = A - np.outer(B, g)
        h = p.t_s
        lhs = np.eye(5) - h / 2 * Ac
        rhs = (np.eye(5) + h / 2 * Ac) @ x[:5] + h * B * g[4] * psi_c
        xb = np.linalg.solve(lhs, rhs)
        pm = (x[4] + xb[4]) / 2
        px = x[5] + h * vbar * math.cos(pm)
        py = x[6] + h * vbar * math.sin(pm)
        self.x[k] = np.r_[xb, px, py]
        self.v[k] = v
        xn = self.x[k]
        self.s[k] = (xn[5], -xn[6], -limit_angle(xn[4]), v, -limit_angle(xn[1]),
                     limit_angle(xn[0]), -xn[3], xn[2])                # :347-356
        self._record(k, wrap=False)                                   # vehicle.py:320-321

    def step_planarpoint(self, k, Fx, Fy):
        """PlanarPointDynamics.step, dynamics.py:1051-1079 (closed-form midpoint)."""
        p = self.p
        vd = math.sqrt(Fx**2 + Fy**2)
        a = thresh(p.k_p_v * (vd - self.v[k]), p.a_max)
        v = thresh(self.v[k] + p.t_s * a, p.v_max_riding)
        vbar = (v + self.s[k, 3]) / 2
        psi_c = limit_angle(math.atan2(Fy, Fx))                        # :112-121
        h, kp = p.t_s, p.k_psi
        x = self.x[k]
        psi_n = ((1 - h * kp / 2) * x[0] + h * kp * psi_c) / (1 + h * kp / 2)
        pm = (x[0] + psi_n) / 2
        self.x[k] = (psi_n, x[1] + h * vbar * math.cos(pm), x[2] + h * vbar * math.sin(pm))
        self.v[k] = v
        self.s[k] = (self.x[k, 1], self.x[k, 2], limit_angle(psi_n), v)  # :959-964
        self._record(k, wrap=False)

    def step_uncontrolled(self, k, Fx, Fy):
        """UncontrolledVehicle.step, vehicle.py:961-979."""
        self.i[k] += 1
        if self.utraj is not None and self.utraj[k] is not None:
            tr = np.asarray(self.utraj[k])
            if tr.shape[1] > self.i[k]:
                self.s[k] = tr[:, self.i[k]]

    def step_agent(self, k, Fx, Fy):
        getattr(self, "step_" + self.model)(k, Fx, Fy)


def invpend_gains(v):
    """InvPendulumBicycleParameters.fullstate_feedback_gains, parameters.py:1857-1892."""
    params_kx = np.array([
        [3.48203226e02, -5.12057324e03, 1.58364873e04, -1.98073306e04],
        [-4.51700000e01, 0.0, 0.0, 0.0],
        [-9.16379250e02, 1.31769807e04, -6.57341643e04, 8.22163589e04],
        [3.20214069e02, -4.69953797e03, 1.66378680e04, -2.43114309e04],
        [2.87549256e-08, -2.27913445e03, 0.0, 0.0],
    ])
    params_ku = np.array([-3.38638984e-09, -2.27913445e03, 0.0, 0.0])
    vdata = np.array((1, v**-1, v**-2, v**-3))
    return params_kx @ vdata, params_ku @ vdata


# ---------------------------------------------------------------------------
# the intersection  (a1, a3)
# ---------------------------------------------------------------------------
class World:
    """SocialForceIntersection restated: groups of Agents + road edges."""

    def __init__(self, groups, priority_rule="unregulated", road_edges=()):
        self.groups = list(groups) if isinstance(groups, (list, tuple)) else [groups]
        self.p2r = priority_rule == "p2r"
        self.road_edges = list(road_edges)      # [(vertices (M,2), F_0, sigma)]

    @property
    def n(self):
        return sum(g.n for g in self.groups)

    def _xypsi(self):
        x = np.concatenate([g.s[:, 0] for g in self.groups])
        y = np.concatenate([g.s[:, 1] for g in self.groups])
        psi = np.concatenate([g.s[:, 2] for g in self.groups])
        return x, y, psi

    def _field(self):
        fp = np.concatenate([np.broadcast_to(field_params_array([g.p])[0], (g.n, 8))
                             for g in self.groups])
        kind = np.concatenate([np.full(g.n, 1 if g.model == "bicycle" else 0) for g in self.groups])
        v = np.concatenate([g.s[:, 3] for g in self.groups])
        bp = next((g.p for g in self.groups if g.model == "bicycle"), None)
        return fp, kind, v, bp

    def calc_forces(self):
        """SocialForceIntersection.calc_forces, intersection.py:747-864."""
        n = self.n
        x, y, psi = self._xypsi()
        fd = np.zeros((n, 2))
        o = 0
        for g in self.groups:
            for k in range(g.n):
                fd[o + k] = g.calc_destination_force(k)
            o += g.n
        if n > 1:
            fp, kind, v, bp = self._field()
            fr = pair_forces(x, y, psi, fp, p2r=self.p2r, src_kind=kind, src_v=v,
                             bicycle_params=bp)
            frx, fry = limit_magnitude(fr[:, 0], fr[:, 1], np.sqrt(fd[:, 0] ** 2 + fd[:, 1] ** 2))
            fr = np.c_[frx, fry]
            F = fr + fd
        else:
            fr = np.zeros((n, 2))
            F = fd.copy()
        for verts, F_0, sigma in self.road_edges:
            fx, fy = road_forces(x, y, verts, F_0, sigma)
            F[:, 0] += fx
            F[:, 1] += fy
        o = 0
        for g in self.groups:
            g.force = F[o:o + g.n].copy()
            g.fdest = fd[o:o + g.n].copy()
            g.frep = fr[o:o + g.n].copy()
            o += g.n
        return F[:, 0].copy(), F[:, 1].copy()

    def step(self):
        """SocialForceIntersection.step, intersection.py:866-896."""
        if self.n > 0:
            self.calc_forces()
            for g in self.groups:
                for k in range(g.n):
                    g.step_agent(k, g.force[k, 0], g.force[k, 1])

    def states(self):
        return [g.s.copy() for g in self.groups]


# ---------------------------------------------------------------------------
# synthetic crowd (SURVEY 8d recipe)
# ---------------------------------------------------------------------------
def synthetic_crowd(n, seed=1, spacing=4.0, n_dest=5, dest_step=60.0, v0=5.0, n_states=5):
    """Seeded open-plane crowd: side L = spacing*sqrt(N); x,y ~ U(0,L); psi ~ U(-pi,pi);
    v = v0; destinations every ``dest_step`` m along psi + U(-0.5, 0.5), stop = 0."""
    rng = np.random.default_rng(seed)
    L = spacing * math.sqrt(n)
    x = rng.uniform(0, L, n)
    y = rng.uniform(0, L, n)
    psi = rng.uniform(-np.pi, np.pi, n)
    s0 = np.zeros((n, n_states))
    s0[:, 0], s0[:, 1], s0[:, 2], s0[:, 3] = x, y, psi, v0
    a = psi + rng.uniform(-0.5, 0.5, n)
    d = dest_step * np.arange(1, n_dest + 1)
    q = np.zeros((n, n_dest, 3))
    q[:, :, 0] = x[:, None] + d[None, :] * np.cos(a)[:, None]
    q[:, :, 1] = y[:, None] + d[None, :] * np.sin(a)[:, None]
    return s0, q
