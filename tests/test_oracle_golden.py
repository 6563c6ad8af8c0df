"""The restated oracle reproduces the vectors captured from the reference's own code
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import csf_oracle as co
from helpers import oracle_world, run_oracle

TOL = 5e-11      # absolute, on O(1..50) quantities; observed <= 7e-13


@pytest.mark.parametrize("model", ["twod", "planarpoint", "bicycle", "invpendulum", "balancingrider"])
def test_demo_geometry(golden, model):
    g = golden
    steps = int(g[f"demo_{model}_steps"].max())
    W = oracle_world(model, g["demo_s0"], g["demo_vd"], g["demo_dests"])
    S, F = run_oracle(W, steps, set(g[f"demo_{model}_steps"].tolist()))
    assert np.abs(S - g[f"demo_{model}_s"]).max() < TOL
    assert np.abs(F - g[f"demo_{model}_F"]).max() < TOL
    assert np.array_equal(W.groups[0].ptr, g[f"demo_{model}_ptr"])
    assert np.array_equal(W.groups[0].znav, g[f"demo_{model}_znav"])


@pytest.mark.parametrize("model", ["twod", "invpendulum", "bicycle"])
def test_stop_destinations(golden, model):
    g = golden
    W = oracle_world(model, g["demo_s0"], g["demo_vd"], g["stop_dests"])
    S, F = run_oracle(W, 2200, set(g[f"stop_{model}_steps"].tolist()))
    assert np.abs(S - g[f"stop_{model}_s"]).max() < TOL
    assert np.abs(F - g[f"stop_{model}_F"]).max() < TOL
    assert np.array_equal(W.groups[0].znav, g[f"stop_{model}_znav"])
    if model != "invpendulum":
        # every bike has come to a halt in state "arrived" (the inverted-pendulum bikes
        # keep circling the stop at walking speed > v_max_stop in the reference, too)
        assert W.groups[0].znav[:, 2].all()
        assert np.all(np.abs(W.groups[0].s[:, 3]) < 1e-12)


def test_priority_to_the_right(golden):
    g = golden
    W = oracle_world("twod", g["demo_s0"], g["demo_vd"], g["demo_dests"], priority_rule="p2r")
    S, F = run_oracle(W, 700, set(g["p2r_twod_steps"].tolist()))
    assert np.abs(S - g["p2r_twod_s"]).max() < TOL
    assert np.abs(F - g["p2r_twod_F"]).max() < TOL


def test_synthetic_crowd(golden):
    g = golden
    W = oracle_world("twod", g["crowd_s0"], np.full(24, 5.0), g["crowd_dests"])
    S, F = run_oracle(W, 60, set(g["crowd_twod_steps"].tolist()))
    assert np.abs(S - g["crowd_twod_s"]).max() < TOL
    assert np.abs(F - g["crowd_twod_F"]).max() < TOL


def test_pair_matrix_and_mask(golden):
    g = golden
    x, y, psi = g["crowd_pair_xypsi"].T
    p = co.default_params("twod")
    fp = co.field_params_array([p])[0]
    tr = co.tracked_mask(x, y, psi, p.hfov)
    assert np.array_equal(tr, g["crowd_pair_tracked"])
    # SURVEY Appendix B counted 205 of 552 pairs at step 5; this is the step-60 state
    assert 0 < tr.sum() < 552
    Fx, Fy = co.twod_field(x[:, None], y[:, None], psi[:, None], fp[None, None, :],
                           x[None, :], y[None, :], psi[None, :])
    Fx = np.where(tr, Fx, 0.0)
    Fy = np.where(tr, Fy, 0.0)
    assert np.abs(Fx - g["crowd_pair_Fx"]).max() < 1e-13
    assert np.abs(Fy - g["crowd_pair_Fy"]).max() < 1e-13
    fr = co.pair_forces(x, y, psi, fp)
    assert np.abs(fr[:, 0] - g["crowd_pair_Fx"].sum(0)).max() < 1e-12
    # |F_pair| == P exactly (vehicle.py:1644-1646): bounded by f_0
    assert np.hypot(Fx, Fy).max() <= p.f_0 + 1e-12


def test_road_force(golden):
    g = golden
    fx = np.zeros(16)
    fy = np.zeros(16)
    for k in range(4):
        a, b = co.road_forces(g["road_pts"][:, 0], g["road_pts"][:, 1], g[f"road_edge{k}"], 0.15, 2.0)
        fx += a
        fy += b
    assert np.abs(np.c_[fx, fy] - g["road_F"]).max() < 1e-12


def test_utils(golden):
    g = golden
    la = np.array([co.limit_angle(float(a)) for a in g["util_angles"]])
    assert np.array_equal(la, g["util_limit"])
    assert np.array_equal(co.limit_angle(g["util_angles"].copy()), g["util_limit"])
    ad = np.array([co.angle_difference(float(a), float(b)) for a, b in zip(g["util_a1"], g["util_a2"])])
    assert np.array_equal(ad, g["util_angdiff"])
    assert np.array_equal(co.angle_difference(g["util_a1"].copy(), g["util_a2"].copy()), g["util_angdiff"])
    assert (la > -np.pi - 1e-15).all() and (la <= np.pi).all()


def test_invpendulum_gains_and_closed_loop(golden):
    g = golden
    kx, ku = co.invpend_gains(5.0)
    assert np.allclose(kx, g["invpend_kx_v5"], rtol=0, atol=1e-12)
    assert abs(ku - g["invpend_ku_v5"]) < 1e-12
    # closed-loop eigenvalues at v = 5
    A = co.Agents("invpendulum", np.array([[0, 0, 0, 5.0, 0, 0]]))
    Ac, Bc = A.invpend_closed_loop(5.0)
    ev = np.sort_complex(np.linalg.eigvals(Ac))
    ref = np.sort_complex(np.array([-25.537 + 24.242j, -25.537 - 24.242j, -13.140,
                                    -2.393 + 3.629j, -2.393 - 3.629j]))
    assert np.abs(ev - ref).max() < 2e-3
    assert abs(A.p.tau_1_squared - 0.10577993368249616) < 1e-16


def test_balancingrider_matrices(golden):
    g = golden
    M, C1, K0, K2 = co.meijaard_canonical(co.BALANCEASSIST)
    for a, b in ((M, "br_M"), (C1, "br_C1"), (K0, "br_K0"), (K2, "br_K2")):
        assert np.abs(a - g[b]).max() < 1e-13
    A, B = co.balancingrider_matrices(co.BALANCEASSIST, 5.0)
    assert np.abs(A - g["br_A_v5"]).max() < 1e-12
    assert np.abs(B - g["br_B_v5"]).max() < 1e-13
    poles = co.balancingrider_poles(("BR1", 0), 5.0)
    assert np.abs(np.sort_complex(poles) - np.sort_complex(g["br_poles_v5"])).max() < 1e-12
    K = co.place_gain(A, B, poles)
    assert np.abs(K - g["br_gains_v5"]).max() < 1e-9
    # closed loop really has the requested poles
    ev = np.linalg.eigvals(A - np.outer(B, K))
    assert np.abs(np.sort_complex(ev) - np.sort_complex(poles)).max() < 1e-8


def test_meijaard_benchmark_bicycle():
    """External KAT for the matrix builder: the benchmark bicycle of Meijaard et al.
    (2007), Table 1 -> eqs. (5.x) canonical matrices as printed in the paper."""
    bm = dict(w=1.02, c=0.08, lam=np.pi / 10, g=9.81, rR=0.3, mR=2.0, IRxx=0.0603, IRyy=0.12,
              xB=0.3, zB=-0.9, mB=85.0, IBxx=9.2, IBxz=2.4, IByy=11.0, IBzz=2.8,
              xH=0.9, zH=-0.7, mH=4.0, IHxx=0.05892, IHxz=-0.00756, IHyy=0.06, IHzz=0.00708,
              rF=0.35, mF=3.0, IFxx=0.1405, IFyy=0.28)
    M, C1, K0, K2 = co.meijaard_canonical(bm)
    assert np.allclose(M, [[80.81722, 2.31941332208709], [2.31941332208709, 0.29784188199686]],
                       rtol=0, atol=1e-12)
    assert np.allclose(K0, [[-80.95, -2.59951685249872], [-2.59951685249872, -0.80329488458618]],
                       rtol=0, atol=1e-12)
    assert np.allclose(K2, [[0, 76.59734589573222], [0, 2.65431523794604]], rtol=0, atol=1e-12)
    assert np.allclose(C1, [[0, 33.86641391492494], [-0.85035641456978, 1.68540397397560]],
                       rtol=0, atol=1e-12)


def test_parcours(golden):
    g = golden
    A = co.Agents("balancingrider", np.array([[0, 0, np.pi / 2, 5, 0, 0, 0, 0]]), v_desired=[4.0])
    A.set_destinations(0, g["parcours_dests"][:, 0], g["parcours_dests"][:, 1])
    W = co.World([A])
    S, F = run_oracle(W, 1500, set(g["parcours_steps"].tolist()))
    assert np.abs(S - g["parcours_s"]).max() < 1e-9
    assert np.abs(F - g["parcours_F"]).max() < 1e-9
    assert np.array_equal(A.ptr, g["parcours_ptr"])


def test_survey_appendix_b_kats(golden):
    """SURVEY.md Appendix B (captured independently during the survey)."""
    g = golden
    s = g["demo_twod_s"]
    steps = g["demo_twod_steps"].tolist()
    assert np.allclose(s[steps.index(1)][0], [-5.9503, 5.050240813739e-08, 1.016145032946e-06,
                                              4.97, 2.044557410069e-05], rtol=1e-9, atol=1e-15)
    assert np.allclose(s[steps.index(700)][0], [21.36675637402, -0.5206918408874, 4.191649577657e-02,
                                                4.499998378795, -7.407037850475e-04], rtol=1e-9)
    f1 = g["demo_twod_F"][steps.index(1)].ravel()
    assert np.allclose(f1, [4.494503087116, 9.189269719382e-04, 4.580390582814e-04, 4.997545117860,
                            3.432309238526e-04, 4.997244378300], rtol=1e-9)
    pp = g["demo_planarpoint_s"][g["demo_planarpoint_steps"].tolist().index(700)]
    assert np.allclose(pp[0], [22.903499079388, -0.669743650247, 0.05499566999, 4.5], rtol=1e-8)


def test_reference_far_field_nan_quirk():
    """The reference's normalisation underflows for far pairs (vehicle.py:1644-1646): NaN.
    The oracle's default (and the CUDA kernels) return P * unit vector instead."""
    p = co.default_params("twod")
    fp = co.field_params_array([p])[0]
    x = np.array([0.0, -100.0])       # target 141 m behind/abeam of the source
    y = np.array([0.0, 100.0])
    psi = np.array([0.0, 0.1])
    lit = co.twod_field(x[0], y[0], psi[0], fp, x[1], y[1], psi[1], literal=True)
    rob = co.twod_field(x[0], y[0], psi[0], fp, x[1], y[1], psi[1])
    assert np.isnan(lit[0]) and np.isnan(lit[1])
    assert np.isfinite(rob[0]) and np.isfinite(rob[1]) and abs(rob[0]) < 1e-100
    # where the reference is finite the two agree to rounding
    x2, y2 = np.array([3.0]), np.array([1.0])
    a = co.twod_field(0.0, 0.0, 0.3, fp, x2, y2, np.array([2.0]), literal=True)
    b = co.twod_field(0.0, 0.0, 0.3, fp, x2, y2, np.array([2.0]))
    assert abs(a[0] - b[0]).max() < 1e-15 and abs(a[1] - b[1]).max() < 1e-15
