"""The far-field bounds the f32 pair kernel culls with are rigorous: host-only entry points of the
C-ABI library (no GPU needed) against dense numpy sampling of the oracle's field."""
import ctypes as C

import numpy as np
import pytest

from cyclistsocialforce_b200 import _lib, parameters as P
from oracle import csf_oracle as co

LN2 = float(np.log(2.0))


def _rate(p, absphi, s2):
    """decay rate q / sigma of |F| = f_0 exp(-rho q / sigma) (reference vehicle.py:1596-1613)."""
    c = np.cos(absphi)
    A = p.sigma_0 + p.sigma_1 * s2
    B = p.sigma_2 + p.sigma_3 * s2
    e = p.e_0 - p.e_1 * s2
    sig = A - B * np.sqrt((1.0 - c) / 2.0)
    return np.sqrt(1.0 - (e * c) ** 2) / sig


@pytest.mark.parametrize("over", [dict(), dict(hfov=1.2 * np.pi, f_0=5.0, sigma_1=6.0, e_0=0.9)])
def test_cutoff_and_reach_table_are_conservative(over):
    lib = _lib.load()
    par = P.InvPendulumBicycleParameters(**over)
    fp = par.to_field_params(2.0 ** -18, False)
    p = co.default_params("twod", **over)
    phi = np.linspace(0.0, np.pi, 4001)
    s2 = np.linspace(0.0, 1.0, 1001)
    R = _rate(p, phi[:, None], s2[None, :])
    dcut = lib.csf_field_cutoff_distance(C.byref(fp))
    true_cut = 40.0 * LN2 / R.min()
    assert true_cut <= dcut <= 1.03 * true_cut            # rigorous (>=) and tight (3 %)
    if not over:
        assert abs(dcut - 159.7) < 0.5
    # reach table: bin b must cover every direction with cos(phi) <= upper edge of the bin
    tab = (C.c_double * 64)()
    assert lib.csf_field_reach_table(C.byref(fp), 64, tab) == 0
    tab = np.array(tab)
    reach = 40.0 * LN2 / R.min(axis=1)                     # per direction, worst heading difference
    assert np.all(np.diff(tab) >= 0) and abs(tab[-1] - dcut) < 2e-4 * dcut   # (1.0001 safety factor)
    for b in range(64):
        upper = -1.0 + 2.0 * (b + 1) / 64
        sel = np.cos(phi) <= upper + 1e-12
        assert tab[b] >= reach[sel].max() * (1 - 1e-9), b
    # tightness: the table is a step function of the true (monotone envelope of the) reach
    env = np.maximum.accumulate(reach[::-1])[::-1]
    assert tab[32] <= 1.2 * env[np.searchsorted(-np.cos(phi), -(-1.0 + 2.0 * 33 / 64))]
    # a larger cut-off exponent reaches further
    fp2 = par.to_field_params(2.0 ** -18, False)
    fp2.cutoff_log2 = 50.0
    assert lib.csf_field_cutoff_distance(C.byref(fp2)) > dcut


def test_lobe_filter_never_drops_a_relevant_pair():
    """numpy replica of the kernel's filter (lobe_reaches in csf_pair_tiled.cu) on random
    source / target-block configurations: whenever it drops a source, the oracle's force of that
    source on every target of the block is below 2^-40 f_0."""
    lib = _lib.load()
    par = P.InvPendulumBicycleParameters()
    fp = par.to_field_params(2.0 ** -18, False)
    tab = (C.c_double * 64)()
    lib.csf_field_reach_table(C.byref(fp), 64, tab)
    tab = np.array(tab)
    p = co.default_params("twod")
    fpar = co.field_params_array([p])[0]
    rng = np.random.default_rng(5)
    n_src, n_tgt = 4000, 24
    dropped = kept = 0
    worst = 0.0
    for trial in range(40):
        Rb = rng.uniform(2.0, 40.0)
        ang = rng.uniform(0, 2 * np.pi, n_tgt)
        rad = Rb * np.sqrt(rng.uniform(0, 1, n_tgt))
        tx, ty = rad * np.cos(ang), rad * np.sin(ang)          # targets inside the block circle (centre 0,0)
        tpsi = rng.uniform(-np.pi, np.pi, n_tgt)
        d = rng.uniform(0.0, 260.0, n_src)
        a = rng.uniform(0, 2 * np.pi, n_src)
        sx, sy, spsi = d * np.cos(a), d * np.sin(a), rng.uniform(-np.pi, np.pi, n_src)
        # the filter (block centre - source)
        dx, dy = -sx, -sy
        dist = np.hypot(dx, dy) + 1e-300
        cphi = (dx * np.cos(spsi) + dy * np.sin(spsi)) / dist
        sphi = np.abs(dy * np.cos(spsi) - dx * np.sin(spsi)) / dist
        sdel = np.minimum(Rb / dist, 1.0)
        cdel = np.sqrt(np.maximum(1.0 - sdel ** 2, 0.0))
        cmin = np.where(cphi >= cdel, 1.0, cphi * cdel + sphi * sdel)
        b = np.clip(((cmin + 1.00002) * 32).astype(int), 0, 63)
        keep = dist - Rb <= tab[b] * 1.0001
        # oracle: field of every source at every target, no mask (the mask only removes more)
        Fx, Fy = co.twod_field(sx[:, None], sy[:, None], spsi[:, None], fpar[None, None, :],
                               tx[None, :], ty[None, :], tpsi[None, :])
        mag = np.hypot(Fx, Fy).max(axis=1) / p.f_0
        if (~keep).any():
            worst = max(worst, float(mag[~keep].max()))
        dropped += int((~keep).sum())
        kept += int(keep.sum())
    assert dropped > 20000 and kept > 20000
    assert worst < 2.0 ** -40, worst


# ---- v0.1 Bicycle elliptic field (reference vehicle.py:1054-1147) in the tiled kernel ---------------------------
def _bike_fp(cutoff=None):
    par = P.BicycleParameters()
    fp = par.to_field_params(2.0 ** -18, False, 1)
    if cutoff is not None:
        fp.cutoff_log2 = cutoff
    return par, fp


def test_bicycle_field_cutoff_is_conservative():
    """Beyond csf_field_cutoff_distance (field_kind 1) the oracle's Bicycle field is below 2^-40 p_0 / p_decay for
    every eccentricity the model can take (e <= 0.7) and every direction; the bound is tight straight ahead."""
    lib = _lib.load()
    par, fp = _bike_fp()
    p = co.default_params("bicycle")
    amp = p.p_0 / p.p_decay
    dcut = lib.csf_field_cutoff_distance(C.byref(fp))
    assert abs(dcut - 41.5 * LN2 * p.p_decay * np.sqrt(1.7 / 0.3)) < 1e-9 * dcut
    v = np.r_[0.0, np.geomspace(1e-6, 1.0, 60) * p.v_max_riding[1], 3.0 * p.v_max_riding[1]]   # e from 0 to the 0.7 cap
    phi = np.linspace(-np.pi, np.pi, 721)
    V, PHI = np.meshgrid(v, phi, indexing="ij")
    for scale, below in ((1.0, True), (0.9, False)):
        Fx, Fy = co.bicycle_field(0.0, 0.0, 0.0, V, p, scale * dcut * np.cos(PHI), scale * dcut * np.sin(PHI))
        worst = np.hypot(Fx, Fy).max() / amp
        assert (worst < 2.0 ** -40) == below, (scale, worst)
    # the reach table is the level-set ellipse at e = 0.7: conservative per direction
    tab = (C.c_double * 64)()
    assert lib.csf_field_reach_table(C.byref(fp), 64, tab) == 0
    tab = np.array(tab)
    assert np.all(np.diff(tab) > 0) and abs(tab[-1] - dcut) < 1e-9 * dcut
    vcap = np.full_like(phi, p.v_max_riding[1])
    for b in range(64):
        upper = -1.0 + 2.0 * (b + 1) / 64
        sel = np.cos(phi) <= upper + 1e-12
        Fx, Fy = co.bicycle_field(0.0, 0.0, 0.0, vcap[sel], p, tab[b] * np.cos(phi[sel]), tab[b] * np.sin(phi[sel]))
        assert np.hypot(Fx, Fy).max() / amp < 2.0 ** -40, b
    fp2 = _bike_fp(50.0)[1]
    assert lib.csf_field_cutoff_distance(C.byref(fp2)) > dcut


def test_bicycle_filter_never_drops_a_relevant_pair():
    """numpy replica of the kernel's source filter for the Bicycle field (bike_reaches in csf_pair_tiled.cu; the
    sorted copy carries the heading scaled by the eccentricity) on random source / target-block configurations:
    whenever it drops a source, the oracle's force of that source on every target inside the block circle is
    below 2^-40 p_0 / p_decay -- and it does drop most of the sources that are out of reach."""
    lib = _lib.load()
    par, fp = _bike_fp()
    p = co.default_params("bicycle")
    amp = p.p_0 / p.p_decay
    K = lib.csf_field_cutoff_distance(C.byref(fp)) / np.sqrt(1.7 / 0.3)
    rng = np.random.default_rng(11)
    n_src, n_tgt = 4000, 24
    dropped = kept = useless_kept = 0
    worst = 0.0
    for trial in range(40):
        Rb = rng.uniform(2.0, 40.0)
        ang = rng.uniform(0, 2 * np.pi, n_tgt)
        rad = Rb * np.sqrt(rng.uniform(0, 1, n_tgt))
        rad[:4] = Rb                                               # targets on the rim too
        tx, ty = rad * np.cos(ang), rad * np.sin(ang)
        d = rng.uniform(0.0, 420.0, n_src)
        a = rng.uniform(0, 2 * np.pi, n_src)
        sx, sy, spsi = d * np.cos(a), d * np.sin(a), rng.uniform(-np.pi, np.pi, n_src)
        sv = np.where(rng.uniform(size=n_src) < 0.3, rng.uniform(0, 1e-3, n_src), rng.uniform(0, 9.0, n_src))
        e = np.minimum(np.power(sv / p.v_max_riding[1], 0.1), 0.7)
        c, s = e * np.cos(spsi), e * np.sin(spsi)                 # what the sorted copy holds
        dx, dy = -sx, -sy                                         # block centre - source
        dist = np.hypot(dx, dy) + 1e-300
        e2 = np.minimum(c * c + s * s, 0.5)
        ee = np.sqrt(e2)
        cphi = (dx * c + dy * s) / dist
        sphi = np.abs(dy * c - dx * s) / dist
        sdel = np.minimum(Rb / dist, 1.0)
        cdel = np.sqrt(np.maximum(1.0 - sdel ** 2, 0.0))
        w = np.minimum(np.where(cphi >= ee * cdel, ee, cphi * cdel + sphi * sdel), ee)
        keep = (dist - Rb) * (1.0 - w) <= K * np.sqrt(1.0 - e2) * 1.0001
        Fx, Fy = co.bicycle_field(sx[:, None], sy[:, None], spsi[:, None], sv[:, None], p, tx[None, :], ty[None, :])
        mag = np.hypot(Fx, Fy).max(axis=1) / amp
        if (~keep).any():
            worst = max(worst, float(mag[~keep].max()))
        dropped += int((~keep).sum())
        kept += int(keep.sum())
        useless_kept += int((keep & (mag < 2.0 ** -60)).sum())
    assert dropped > 20000 and kept > 20000
    assert worst < 2.0 ** -40, worst
    assert useless_kept < 0.2 * kept                               # the filter is not vacuous
