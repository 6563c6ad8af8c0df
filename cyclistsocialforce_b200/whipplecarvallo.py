"""Linearised Whipple-Carvallo bicycle (Meijaard, Papadopoulos, Ruina & Schwab 2007).

Host-side constant builder for the BalancingRiderBicycle kernel.  The reference
obtains these matrices from the third-party ``bicycleparameters`` package
(reference parameters.py:1285-1286, dynamics.py:522, :572); that package is not a
dependency here, the canonical matrices are formed directly from the paper's
Appendix A.
"""
from __future__ import annotations

import math

import numpy as np

#: reference data/bicycleparams/balanceassist_bikeparams.py:11-39 (BSD-2 data set of
#: the BicycleParameters project: Balanceassistv1 bicycle + rider "Jason")
balanceassistv1_with_averagerider = dict(
    IBxx=16.136560964517308, IBxz=-2.5375819134691833, IByy=18.98228436804581,
    IBzz=4.308368614306412, IFxx=0.0995, IFyy=0.1902, IHxx=0.2984, IHxz=-0.038,
    IHyy=0.257, IHzz=0.0566, IRxx=0.1023, IRyy=0.1887, c=0.042, g=9.81, lam=0.255,
    mB=91.50000000000003, mF=2.235, mH=4.3, mR=4.085, rF=0.35231, rR=0.34895, v=1.0,
    w=1.113, xB=0.373106714751133, xH=0.921, yB=0.0, zB=-0.9697039390081493, zH=-0.86,
)


def canonical_matrices(b: dict):
    """(M, C1, K0, K2) with  M q'' + v C1 q' + (g K0 + v^2 K2) q = f,  q = [roll, steer]."""
    sl, cl = math.sin(b["lam"]), math.cos(b["lam"])
    w, c = b["w"], b["c"]
    # total system
    m_t = b["mR"] + b["mB"] + b["mH"] + b["mF"]
    x_t = (b["xB"] * b["mB"] + b["xH"] * b["mH"] + w * b["mF"]) / m_t
    z_t = (-b["rR"] * b["mR"] + b["zB"] * b["mB"] + b["zH"] * b["mH"] - b["rF"] * b["mF"]) / m_t
    i_txx = (b["IRxx"] + b["IBxx"] + b["IHxx"] + b["IFxx"] + b["mR"] * b["rR"] ** 2
             + b["mB"] * b["zB"] ** 2 + b["mH"] * b["zH"] ** 2 + b["mF"] * b["rF"] ** 2)
    i_txz = (b["IBxz"] + b["IHxz"] - b["mB"] * b["xB"] * b["zB"] - b["mH"] * b["xH"] * b["zH"]
             + b["mF"] * w * b["rF"])
    i_tzz = (b["IRxx"] + b["IBzz"] + b["IHzz"] + b["IFxx"] + b["mB"] * b["xB"] ** 2
             + b["mH"] * b["xH"] ** 2 + b["mF"] * w ** 2)
    # front assembly
    m_a = b["mH"] + b["mF"]
    x_a = (b["xH"] * b["mH"] + w * b["mF"]) / m_a
    z_a = (b["zH"] * b["mH"] - b["rF"] * b["mF"]) / m_a
    i_axx = b["IHxx"] + b["IFxx"] + b["mH"] * (b["zH"] - z_a) ** 2 + b["mF"] * (b["rF"] + z_a) ** 2
    i_axz = b["IHxz"] - b["mH"] * (b["xH"] - x_a) * (b["zH"] - z_a) + b["mF"] * (w - x_a) * (b["rF"] + z_a)
    i_azz = b["IHzz"] + b["IFxx"] + b["mH"] * (b["xH"] - x_a) ** 2 + b["mF"] * (w - x_a) ** 2
    u_a = (x_a - w - c) * cl - z_a * sl
    i_all = m_a * u_a ** 2 + i_axx * sl ** 2 + 2 * i_axz * sl * cl + i_azz * cl ** 2
    i_alx = -m_a * u_a * z_a + i_axx * sl + i_axz * cl
    i_alz = m_a * u_a * x_a + i_axz * sl + i_azz * cl
    mu = c / w * cl
    s_r, s_f = b["IRyy"] / b["rR"], b["IFyy"] / b["rF"]
    s_t = s_r + s_f
    s_a = m_a * u_a + mu * m_t * x_t
    m01 = i_alx + mu * i_txz
    M = np.array([[i_txx, m01], [m01, i_all + 2 * mu * i_alz + mu ** 2 * i_tzz]])
    K0 = np.array([[m_t * z_t, -s_a], [-s_a, -s_a * sl]])
    K2 = np.array([[0.0, (s_t - m_t * z_t) / w * cl], [0.0, (s_a + s_f * sl) / w * cl]])
    C1 = np.array([[0.0, mu * s_t + s_f * cl + i_txz / w * cl - mu * m_t * z_t],
                   [-(mu * s_t + s_f * cl), i_alz / w * cl + mu * (s_a + i_tzz / w * cl)]])
    return M, C1, K0, K2


def speed_polynomial_state_matrices(b: dict):
    """A(v) = A0 + v A1 + v^2 A2 (5x5, states roll, steer, roll rate, steer rate, yaw) and the
    steer-torque input column B (reference dynamics.py:511-538, :583-588, :296-302)."""
    M, C1, K0, K2 = canonical_matrices(b)
    Minv = np.linalg.inv(M)
    A0, A1, A2 = np.zeros((5, 5)), np.zeros((5, 5)), np.zeros((5, 5))
    A0[0:2, 2:4] = np.eye(2)
    A0[2:4, 0:2] = -Minv @ (b["g"] * K0)
    A2[2:4, 0:2] = -Minv @ K2
    A1[2:4, 2:4] = -Minv @ C1
    cl = math.cos(b["lam"])
    A1[4, 1] = cl / b["w"]
    A0[4, 3] = cl * b["c"] / b["w"]
    B = np.zeros(5)
    B[2:4] = Minv[:, 1]
    return A0, A1, A2, B
