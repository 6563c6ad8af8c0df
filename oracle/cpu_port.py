"""CPU baseline driver for bench.py (test infrastructure; see oracle/__init__.py).

Times the oracle port of the stepping path on the host cores, on a bounded sample of the
bench workload: ``sample`` agents of the N-agent crowd are advanced per CPU step, each
against all N sources (so the per-agent cost is the full one).  Uses the C/OpenMP
restatement (oracle/csf_oracle_c.c) when it has been built, else the numpy oracle.
"""
from __future__ import annotations

import os
import time

import numpy as np

from oracle import csf_oracle as co


def _numpy_sample_step(A, x, y, psi, fp, idx):
    """One oracle step of agents ``idx`` (group ``A`` holds exactly those agents)."""
    fd = np.array([A.calc_destination_force(k) for k in range(A.n)])
    fr = co.pair_forces(x, y, psi, fp, tgt=idx)
    frx, fry = co.limit_magnitude(fr[:, 0], fr[:, 1], np.hypot(fd[:, 0], fd[:, 1]))
    F = np.c_[frx, fry] + fd
    for k in range(A.n):
        A.step_agent(k, F[k, 0], F[k, 1])
    x[idx], y[idx], psi[idx] = A.s[:, 0], A.s[:, 1], A.s[:, 2]


def timed_sample(n_agents, seed, steps, warmup, sample=None):
    s0, q = co.synthetic_crowd(n_agents, seed=seed)
    try:
        from oracle import c_port
        have_c = c_port.available()
    except Exception:
        have_c = False
    if have_c:
        return c_port.timed_sample(s0, q, steps, warmup, sample)
    sample = sample or 256
    idx = np.arange(sample)
    p = co.default_params("twod")
    fp = co.field_params_array([p])[0]
    A = co.Agents("twod", s0[idx])
    for k in range(sample):
        A.set_destinations(k, q[k, :, 0], q[k, :, 1])
    x, y, psi = s0[:, 0].copy(), s0[:, 1].copy(), s0[:, 2].copy()
    for _ in range(warmup):
        _numpy_sample_step(A, x, y, psi, fp, idx)
    t0 = time.perf_counter()
    for _ in range(steps):
        _numpy_sample_step(A, x, y, psi, fp, idx)
    dt = time.perf_counter() - t0
    return {"value": sample * steps / dt, "ms_per_step": dt / steps * 1e3, "cores": 1, "kind": "port",
            "sample_agents": sample, "ms_per_full_step_extrapolated": dt / steps * 1e3 * (n_agents / sample),
            "sample": f"numpy oracle, {sample} of {n_agents} agents stepped per CPU step, each against all "
                      f"{n_agents} sources; {steps} steps; ms_per_step is the measured time of one such sample step"}


def timed_numpy_oracle_full(n_agents=4096, seed=1, steps=2, warmup=1):
    """the restated, vectorised fp64 numpy oracle (validated against the
    reference's own code at N <= 64) stepping a FULL ``n_agents`` crowd -- every agent, every pair -- on
    one host core.  (The reference's own Python cannot: its mask code is O(N^4), N ~ 200 exhausts memory.)"""
    s0, q = co.synthetic_crowd(n_agents, seed=seed)
    A = co.Agents("twod", s0)
    for k in range(n_agents):
        A.set_destinations(k, q[k, :, 0], q[k, :, 1])
    W = co.World([A])
    for _ in range(warmup):
        W.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        W.step()
    dt = time.perf_counter() - t0
    return {"n_agents": n_agents, "steps": steps, "ms_per_step": dt / steps * 1e3,
            "agent_steps_per_s": n_agents * steps / dt, "cores": 1,
            "what": "restated numpy fp64 oracle (not the reference's code), full crowd, all pairs"}
