"""ctypes wrapper of the C/OpenMP oracle (oracle/csf_oracle_c.c).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

from oracle import csf_oracle as co

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csf_oracle_c.c")
LIB = os.path.join(HERE, "_build", "libcsf_oracle_c.so")
_lib = None


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-o", LIB + ".tmp", SRC, "-lm"]
    subprocess.check_call(cmd)
    os.replace(LIB + ".tmp", LIB)
    return LIB


def available():
    try:
        load()
        return True
    except Exception:
        return False


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.csf_c_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def param_vector(p):
    return np.array([p.t_s, p.d_arrived_inter, p.d_arrived_stop, p.v_max_stop, p.v_max_harddecel,
                     p.a_max[0], p.a_max[1], p.a_desired_default[0], p.a_desired_default[1],
                     p.v_max_riding[0], p.v_max_riding[1], p.l, p.delta_max, p.k_p_v, p.k_p_delta, p.g], float)


def pair_forces(x, y, psi, fp, tgt=None, p2r=False):
    lib = load()
    x, y, psi = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, psi))
    n = x.shape[0]
    tgt = np.arange(n, dtype=np.int64) if tgt is None else np.ascontiguousarray(tgt, dtype=np.int64)
    out = np.zeros((tgt.shape[0], 2))
    fp = np.ascontiguousarray(fp, dtype=np.float64)
    lib.csf_c_pair_forces(C.c_int64(n), _p(x), _p(y), _p(psi), _p(fp), C.c_int(int(p2r)),
                          C.c_int64(tgt.shape[0]), _p(tgt), _p(out))
    return out


class TwoDCrowdC:
    """A TwoDBicycle crowd of n agents of which the first ``ns`` are stepped ("sample")."""

    def __init__(self, s0, dests, ns=None, params=None, vd=None):
        self.p = params or co.default_params("twod")
        self.pv = param_vector(self.p)
        self.fp = co.field_params_array([self.p])[0]
        n = s0.shape[0]
        self.n = n
        self.ns = ns = n if ns is None else ns
        self.x, self.y = s0[:, 0].copy(), s0[:, 1].copy()
        self.psi = np.array([co.limit_angle(float(a)) for a in s0[:, 2]])
        self.s = np.ascontiguousarray(s0[:ns, :5].copy())
        self.s[:, 2] = self.psi[:ns]
        q = co.np.zeros((ns, dests.shape[1] + 1, 3))
        q[:, 0, 0], q[:, 0, 1] = s0[:ns, 0], s0[:ns, 1]
        q[:, 1:, :dests.shape[2]] = dests[:ns]
        self.destq = np.ascontiguousarray(q)
        self.qcap = q.shape[1]
        self.qlen = np.full(ns, q.shape[1], np.int32)
        self.i = np.zeros(ns, np.int32)
        self.ptr = np.zeros(ns, np.int32)
        self.znav = np.ones(ns, np.int32)
        self.znavp = np.zeros((ns, 3))
        self.vd = np.full(ns, self.p.v_desired_default if vd is None else vd, float)
        self.prev = np.ascontiguousarray(s0[:ns, :2].copy())
        self.hist = np.zeros((ns, 128, 2))
        self.hist[:, 0, :] = s0[:ns, :2]
        self.hstep = np.zeros(ns, np.int32)
        self.force = np.zeros((ns, 2))
        self.idx = np.arange(ns, dtype=np.int64)

    def step(self):
        lib = load()
        frep = pair_forces(self.x, self.y, self.psi, self.fp, tgt=self.idx)
        lib.csf_c_twod_step(C.c_int64(self.ns), C.c_int64(self.qcap), _p(self.s), _p(self.i), _p(self.ptr),
                            _p(self.qlen), _p(self.destq), _p(self.znav), _p(self.znavp), _p(self.vd),
                            _p(self.prev), _p(self.hist), _p(self.hstep), _p(self.pv), C.c_int64(self.n),
                            _p(frep), _p(self.force))
        self.x[:self.ns], self.y[:self.ns], self.psi[:self.ns] = self.s[:, 0], self.s[:, 1], self.s[:, 2]


def timed_sample(s0, q, steps, warmup, sample=None, budget_s=20.0):
    lib = load()
    n = s0.shape[0]
    lib.csf_c_set_num_threads(C.c_int(os.cpu_count() or 1))     # every host thread, also under torchrun
    cores = lib.csf_c_num_threads()
    if sample is None:
        # size the sample for ~budget_s of CPU work: probe with 64 agents
        probe = TwoDCrowdC(s0, q, ns=min(64, n))
        t0 = time.perf_counter()
        probe.step()
        per_agent = (time.perf_counter() - t0) / probe.ns
        sample = int(max(64, min(n, budget_s / max(per_agent, 1e-9) / max(steps + warmup, 1))))
        sample = min(n, (sample // 64) * 64)
    crowd = TwoDCrowdC(s0, q, ns=sample)
    for _ in range(warmup):
        crowd.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        crowd.step()
    dt = time.perf_counter() - t0
    return {"value": sample * steps / dt, "ms_per_step": dt / steps * 1e3, "cores": cores, "kind": "port",
            "sample_agents": sample, "ms_per_full_step_extrapolated": dt / steps * 1e3 * (n / sample),
            "sample": f"C/OpenMP oracle ({cores} threads), {sample} of {n} agents stepped per CPU step, each "
                      f"against all {n} sources; {steps} steps; ms_per_step is the measured time of one such "
                      f"sample step"}
