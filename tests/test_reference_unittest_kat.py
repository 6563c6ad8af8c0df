"""The reference's only unit test, restated: src/cyclistsocialforce/test.py:15-119
(``TestBicycleDynamics.test_yaw_stepresponse_invpend``).

That test steps ``InvPendulumBicycle.step_yaw`` 1000 times through a 30 deg yaw step at v = 5 and
compares steer, roll and yaw angle (rtol 1e-7) with a python-control closed loop built by
``ct.place`` with the poles (-0.2, -0.1 +- 0.1j, -0.15, -0.1) * 30 and ``Ku = 1 / -0.46313878281084603``.
It cannot run at the reference's HEAD (python-control is absent here, the constructor raises, and
the gain table of parameters.py:1863-1883 has since been re-fitted to other poles), but the constants
it holds pin the two pieces of third-party arithmetic the InvPendulum path rests on:

* ``place``: the gains it produces for these poles on the reference's open-loop matrices are kept, to
  nine digits, as the commented-out ``K_x`` / ``K_u`` of parameters.py:1858-1861, and the test's ``Ku`` is
  the reciprocal DC gain of that closed loop to sixteen digits;
* ``forced_response``: the per-step propagation of that closed loop (what ``step_yaw`` does, with the
  test's gains) against its continuous-time step response, 1000 steps, rtol 1e-7.

Both are checked on the oracle (CPU) and the second one on the f64 build of the kernel (GPU).
"""
import numpy as np
import pytest

from oracle import csf_oracle as co

# constants held by the reference
POLES = np.array((-0.2 + 0j, -0.1 + 0.1j, -0.1 - 0.1j, -0.15 + 0j, -0.1 + 0j)) * 30      # test.py:85-90
KU = 1 / -0.46313878281084603                                                            # test.py:81
KX_COMMENT = np.array([6.26092881, -48.635, -6.92845026, -2.25215286, -2.15918001])      # parameters.py:1858-1860
KU_COMMENT = -2.1591800063357907                                                         # parameters.py:1861
N_STEPS = 1000
I_STEP = 200                                                                             # int(0.2 * len(t)), test.py:33


def open_loop(p, v):
    """test.py:52-77 == vehicle.py:1738-1768."""
    K = v ** 2 / (p.g * p.l)
    tau_2 = p.l_2 / v
    tau_3 = p.l / v
    A = np.array([[0, 1, 0, 0, 0],
                  [0, -p.c_steer / p.i_steer_vertvert, 0, 0, 0],
                  [0, 0, 0, 1, 0],
                  [-K / p.tau_1_squared, -K * tau_2 / p.tau_1_squared, 1 / p.tau_1_squared, 0, 0.0],
                  [1 / tau_3, 0, 0, 0, 0]])
    B = np.array([0, 1 / p.i_steer_vertvert, 0, 0, 0])
    return A, B


def _inputs():
    p = co.default_params("invpendulum")
    t = np.arange(0, 10, p.t_s)
    psi_d = np.zeros_like(t)
    psi_d[int(0.2 * len(t)):] = 2 * np.pi * 30 / 360                                     # test.py:31-33
    return p, t, psi_d


def continuous_step_response(Ac, Bc, t_s, psi_d):
    """States of x' = Ac x + Bc u at t_k = k t_s for the piecewise-constant input u = psi_d[k] on
    [t_k, t_k+1), from zero: after the step at I_STEP, x(tau) = Ac^-1 (e^{Ac tau} - I) Bc u  (one matrix
    exponential per sample from the eigen-decomposition, not a recursion)."""
    w, V = np.linalg.eig(Ac)
    Vi = np.linalg.inv(V)
    u = psi_d[-1]
    x = np.zeros((5, len(psi_d) + 1))
    g = Vi @ Bc * u
    for k in range(I_STEP + 1, len(psi_d) + 1):
        tau = (k - I_STEP) * t_s
        x[:, k] = np.real(V @ (np.expm1(w * tau) / w * g))
    return x


def test_place_reproduces_the_gains_the_reference_kept():
    p, _, _ = _inputs()
    A, B = open_loop(p, p.v_desired_default)
    K = co.place_gain(A, B, POLES)
    assert np.allclose(K, KX_COMMENT, rtol=0, atol=6e-9 * np.abs(KX_COMMENT).max())
    # DC gain u -> psi of ss(A - B K, B, [0 0 0 0 1], 0): the reference's Ku is its reciprocal
    dc = -np.array([0, 0, 0, 0, 1.0]) @ np.linalg.solve(A - np.outer(B, K), B)
    assert abs(dc - (-0.46313878281084603)) < 1e-12
    assert abs(1 / dc - KU_COMMENT) < 1e-10
    assert abs(K[4] - 1 / dc) < 1e-9                     # unit DC gain: k_psi == K_u
    assert np.allclose(np.poly(A - np.outer(B, K)), np.real(np.poly(POLES)), rtol=1e-9)   # the placed poles


def _oracle_yaw_response(monkeypatch):
    p, t, psi_d = _inputs()
    A, B = open_loop(p, p.v_desired_default)
    K = co.place_gain(A, B, POLES)
    monkeypatch.setattr(co, "invpend_gains", lambda v: (K, KU))
    v = p.v_desired_default
    ag = co.Agents("invpendulum", np.array([[0.0, 0.0, 0.0, v, 0.0, 0.0]]))
    traj = np.zeros((3, len(t) + 1))                     # delta, theta, psi  (traj rows 4, 5, 2 of test.py)
    for k in range(len(t)):
        ag.step_invpendulum(0, v * np.cos(psi_d[k]), v * np.sin(psi_d[k]))
        assert ag.s[0, 3] == v                           # "Speed is kept constant."
        traj[:, k + 1] = ag.s[0, 4], ag.s[0, 5], ag.s[0, 2]
    Ac, Bc = A - np.outer(B, K), KU * B
    return traj, continuous_step_response(Ac, Bc, p.t_s, psi_d), (K, KU)


def test_oracle_yaw_step_response_matches_closed_loop(monkeypatch):
    """test.py:102-119 with assert_allclose's default rtol = 1e-7."""
    traj, x, _ = _oracle_yaw_response(monkeypatch)
    assert len(traj[0]) == N_STEPS + 1
    np.testing.assert_allclose(traj[0], x[0], rtol=1e-7, atol=1e-13, err_msg="Error in steer angle!")
    np.testing.assert_allclose(traj[1], x[2], rtol=1e-7, atol=1e-13, err_msg="Error in roll angle!")
    np.testing.assert_allclose(traj[2], x[4], rtol=1e-7, atol=1e-13, err_msg="Error in yaw angle!")
    # the response settles on the commanded yaw (unit DC gain by construction of Ku)
    assert abs(traj[2, -1] - np.deg2rad(30)) < 2e-3


@pytest.mark.gpu
def test_gpu_f64_yaw_step_response_matches_closed_loop():
    """The same response from the f64 build of the per-agent kernel (degree-13 Pade propagator of the
    6x6 augmented closed-loop matrix, csf_agent.cu), gains of the test fed through the gain table."""
    import torch
    from cyclistsocialforce_b200 import parameters as P
    from cyclistsocialforce_b200.engine import AgentGroup, Engine
    p, t, psi_d = _inputs()
    A, B = open_loop(p, p.v_desired_default)
    K = co.place_gain(A, B, POLES)
    params = P.InvPendulumBicycleParameters()
    params.KX_TABLE = tuple((float(k), 0.0, 0.0, 0.0) for k in K)       # gains independent of v
    params.KU_TABLE = (KU, 0.0, 0.0, 0.0)
    v = params.v_desired_default
    n = 3                                                                # same response for every agent
    s0 = np.tile(np.array([[0.0, 0.0, 0.0, v, 0.0, 0.0]]), (n, 1))
    s0[:, 0] = 100.0 * np.arange(n)
    g = AgentGroup("invpendulum", s0, params, dtype=torch.float64)
    eng = Engine([g], dtype=torch.float64)
    traj = np.zeros((3, len(t) + 1))
    for k in range(len(t)):
        eng.force[:, 0] = v * np.cos(psi_d[k])
        eng.force[:, 1] = v * np.sin(psi_d[k])
        eng.advance()                                                    # Vehicle.step(Fx, Fy)
        if k % 100 == 99 or k == len(t) - 1:
            s = g.states_numpy()
            assert np.all(s[:, 3] == v)
            assert np.all(s[:, 2:] == s[0, 2:])
        st = g.states_numpy()[1]
        traj[:, k + 1] = st[4], st[5], st[2]
    eng.check_status()
    x = continuous_step_response(A - np.outer(B, K), KU * B, p.t_s, psi_d)
    np.testing.assert_allclose(traj[0], x[0], rtol=1e-7, atol=1e-13, err_msg="Error in steer angle!")
    np.testing.assert_allclose(traj[1], x[2], rtol=1e-7, atol=1e-13, err_msg="Error in roll angle!")
    np.testing.assert_allclose(traj[2], x[4], rtol=1e-7, atol=1e-13, err_msg="Error in yaw angle!")
