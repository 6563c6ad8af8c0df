"""Shared builders for the parity tests (oracle side)."""
import numpy as np

from oracle import csf_oracle as co


def oracle_world(model, s0, vd, dests, **world_kw):
    A = co.Agents(model, np.asarray(s0)[:, :co.N_STATES[model]], v_desired=vd)
    for k in range(A.n):
        d = np.asarray(dests[k], float)
        A.set_destinations(k, d[:, 0], d[:, 1], d[:, 2] if d.shape[1] > 2 else None)
    return co.World([A], **world_kw)


def run_oracle(W, steps, keep):
    S, F = [], []
    for k in range(1, steps + 1):
        W.step()
        if k in keep:
            S.append(W.groups[0].s.copy())
            F.append(W.groups[0].force.copy())
    return np.array(S), np.array(F)


def relerr(a, b, floor=1e-12):
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)
