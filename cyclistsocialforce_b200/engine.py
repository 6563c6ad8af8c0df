"""Device-resident crowd state and the per-step kernel sequence.

``AgentGroup`` = the agents of one model class as struct-of-arrays torch CUDA
tensors (the layout of ``CsfAgentState`` in include/csf_b200.h).  ``Engine`` owns the
pair payload, force buffers and workspace and issues, per step, through the C ABI:

    K1   csf_pair_forces_*      all-pairs repulsive force (one launch per source class)
         csf_road_forces_*      road-edge force (if any)
    K2+3 csf_agent_step_*       destination force, assembly, control, dynamics, next payload

which is what ``SocialForceIntersection.step()`` (reference intersection.py:866-896)
does with Python loops.  PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import os
import math

import numpy as np
import torch

from . import _lib
from .parameters import VehicleParameters, choose_q_scale, payload_frame

N_STATES = dict(twod=5, invpendulum=6, balancingrider=8, planarpoint=4, bicycle=5, uncontrolled=4)
_WRAP_MODELS = ("twod", "invpendulum", "bicycle")
_STATE_COLS = ("x", "y", "psi", "v", "delta", "theta", "deltadot", "thetadot")
# per-agent device fields and the axis the agent index runs along (churn: AgentGroup.select / concat)
_FIELD_AXIS = dict(vd_default=0, step_i=0, destq=0, dest_len=0, dest_ptr=0, znav=0, znav_v0=0, znav_d0=0, znav_d1=0,
                   prev_x=0, prev_y=0, hist_x=1, hist_y=1, hist_step=0, ip_x=1, ip_zrid=0, ip_delta_run=0,
                   dyn_x=1, dyn_v=0, br_gains=1, br_poles=1, br_vlast=0, br_draws=0, br_stream=0)
HIST_CAP = 128     # smallest ring of past positions (rows); sized per group from t_s, see hist_capacity()


def hist_capacity(params):
    """Rows of the position ring: the look-back of the last-destination spline is int(1 / t_s) steps
    (reference vehicle.py:1486-1492); the kernel indexes the ring modulo a power of two above it."""
    need = int(1 / params.t_s) + 2
    cap = HIST_CAP
    while cap < need:
        cap *= 2
    return cap


def _wrap_angle(a):
    a = np.asarray(a, float)
    a = np.floor(a / (2 * np.pi)) * (-2 * np.pi) + a
    a = np.where(a > np.pi, a - 2 * np.pi, a)
    return np.where(a < -np.pi, a + 2 * np.pi, a)


#: tuning aid: "refresh" (default) re-sorts the pair kernel's items by cost after the first launch that follows a
#: spatial re-sort, "stale" keeps the order computed from the costs of the previous spatial order until the next
#: re-sort (only call 1 refreshes), "none" hands the items out in index order
_ITEM_ORDER_MODE = os.environ.get("CSF_ITEM_ORDER", "refresh")
#: tuning aid: visiting order of a shard's targets -- "partition" (default: the order of the agent numbering) or
#: "curve" (re-sorted along the Hilbert curve of the current bounding box with the sources)
_SHARD_TARGET_ORDER = os.environ.get("CSF_SHARD_TARGET_ORDER", "partition")


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class AgentGroup:
    """Agents of one model class.  ``s0``: (n, >= N_STATES) float64; ``destqueues``: list of
    (Q_k, 3) arrays [x, y, stop] *including* the start position as entry 0 (the reference's
    queue convention, vehicle.py:183-185)."""

    def __init__(self, model, s0, params, vd_default=None, destqueues=None, dtype=torch.float32,
                 device="cuda", q_cap=None, stream_ids=None):
        assert model in N_STATES and model != "uncontrolled"
        _lib.load()
        if not torch.cuda.is_available():
            raise _lib.CsfError("no CUDA device: the csf_b200 engine has no CPU fallback")
        self.model = model
        self.params = params
        self.dtype = dtype
        self.device = torch.device(device)
        ns = N_STATES[model]
        s0 = np.atleast_2d(np.asarray(s0, dtype=np.float64))
        if s0.shape[1] < ns:
            raise ValueError(f"The initial state s0 has to be size {ns}.")
        s0 = s0[:, :ns].copy()
        s0[:, 2] = _wrap_angle(s0[:, 2])                     # vehicle.py:155
        self.n = n = s0.shape[0]
        dev, T, f64, i32 = self.device, dtype, torch.float64, torch.int32

        def tT(a):
            return torch.as_tensor(np.ascontiguousarray(a), dtype=T, device=dev)

        def t64(a):
            return torch.as_tensor(np.ascontiguousarray(a), dtype=f64, device=dev)

        z = np.zeros(n)
        self._make_state_slab(n, {name: (t64(s0[:, k]) if k < 2 else tT(s0[:, k]))
                                  for k, name in enumerate(_STATE_COLS[:ns])})
        vd = np.full(n, getattr(params, "v_desired_default", 0.0)) if vd_default is None else np.broadcast_to(
            np.asarray(vd_default, float), (n,))
        self.vd_default = tT(vd)
        self.step_i = torch.zeros(n, dtype=i32, device=dev)
        # destinations
        if destqueues is None:
            destqueues = [np.array([[s0[k, 0], s0[k, 1], 0.0]]) for k in range(n)]
        lens = np.array([len(q) for q in destqueues], dtype=np.int32)
        self.q_cap = int(max(int(lens.max()) if n else 1, q_cap or 1))
        self.destq_host = np.zeros((n, self.q_cap, 3))
        for k, q in enumerate(destqueues):
            q = np.asarray(q, float)
            self.destq_host[k, :len(q), :q.shape[1]] = q
        self.dest_len_host = lens
        self.destq = t64(self.destq_host)
        self.dest_len = torch.as_tensor(lens, dtype=i32, device=dev)
        self.dest_ptr = torch.zeros(n, dtype=i32, device=dev)
        self.znav = torch.ones(n, dtype=i32, device=dev)          # [go]   vehicle.py:188
        self.znav_v0, self.znav_d0, self.znav_d1 = tT(z), tT(z), tT(z)
        # history (traj[i-1], 1 s look-back ring)
        needs_hist = model in ("twod", "invpendulum", "planarpoint")
        self.prev_x = t64(s0[:, 0]) if needs_hist else None
        self.prev_y = t64(s0[:, 1]) if needs_hist else None
        self.hist_cap = hist_capacity(params)
        if needs_hist:
            self.hist_x = torch.zeros((self.hist_cap, n), dtype=f64, device=dev)
            self.hist_y = torch.zeros((self.hist_cap, n), dtype=f64, device=dev)
            self.hist_x[0] = self.x
            self.hist_y[0] = self.y
            self.hist_step = torch.zeros(n, dtype=i32, device=dev)
        else:
            self.hist_x = self.hist_y = self.hist_step = None
        self.ip_x = self.ip_zrid = self.ip_delta_run = None
        self.dyn_x = self.dyn_v = self.br_gains = None
        self.br_poles = self.br_vlast = self.br_draws = self.br_stream = None
        if model == "invpendulum":
            # vehicle.py:1728-1736
            self.ip_x = t64(np.stack([s0[:, 4], z, s0[:, 5], z, s0[:, 2]]))
            walk = s0[:, 3] < params.v_max_walk
            self.ip_zrid = torch.as_tensor(np.where(walk, 2, 1).astype(np.int32), device=dev)
            ok = np.abs(s0[:, 4]) < params.delta_max_walk
            self.ip_delta_run = torch.as_tensor(ok.astype(np.int32), device=dev)
        elif model == "balancingrider":
            # dynamics.py:361-399 (CSF frame -> bike frame) and :305-306
            self.dyn_x = t64(np.stack([s0[:, 5], -s0[:, 4], s0[:, 7], -s0[:, 6], -s0[:, 2]]))
            self.dyn_v = t64(s0[:, 3])
            self.br_poles = torch.zeros((5, n), dtype=f64, device=dev)
            self.br_vlast = torch.full((n,), -10000.0, dtype=f64, device=dev)     # parameters.py:1311
            self.br_draws = torch.zeros(n, dtype=i32, device=dev)
            # the random stream of a rider is keyed by this number (default: its index in the group); it
            # travels with the rider through select / concat, so churn does not re-key anybody
            ids = np.arange(n, dtype=np.int64) if stream_ids is None else np.asarray(stream_ids, dtype=np.int64).reshape(n)
            self.br_stream = torch.as_tensor(ids, dtype=torch.int64, device=dev)
            fixed = getattr(params, "gains", None)
            if fixed is not None:                      # parameters `gains=`: never re-designed (dynamics.py:606-607)
                self.br_gains = t64(np.repeat(np.asarray(fixed, float).reshape(5, 1), n, axis=1))
            else:
                self.br_gains = torch.zeros((5, n), dtype=f64, device=dev)       # designed on the device below
        elif model == "planarpoint":
            self.dyn_x = t64(s0[:, 2][None, :])
            self.dyn_v = t64(s0[:, 3])
        self._init_status()
        self.payload_offset = 0
        self._cstate = None
        self._cparams = None
        if model == "balancingrider" and getattr(params, "gains", None) is None:
            self.init_gains()

    def init_gains(self):
        """First gains of every BalancingRider (BalancingRiderDynamics.__init__ -> _get_gains(v),
        dynamics.py:305-306, :602-615), on the device: poles from the pole model's regression, the fixed poles,
        or -- stochastic riders -- the first draw of every rider's own stream (``br_stream``); then the
        same Ackermann design the step kernel uses."""
        self._cstate = None
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().csf_br_init(C.byref(self.cstate()), C.byref(self.cparams(1.0)),
                                               C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                       "csf_br_init")
        self._cstate = None

    def _init_status(self):
        """Device status word + its mirror in pinned (host-mapped) memory: the kernels raise the mirror
        whenever they set a status bit, the host polls it every step without a copy or a sync."""
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.status_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.version = 0             # bumped whenever a device pointer or parameter the kernels see changes

    def _make_state_slab(self, n, cols):
        """The CSF state columns (x, y double; psi, v, delta ... in T) are views into ONE device slab, so
        that a host-driven loop moves the whole state with a single copy each way (Engine.step_host).
        ``cols``: name -> device tensor of length n (missing columns become None)."""
        T, f64, dev = self.dtype, torch.float64, self.device
        esz = torch.empty(0, dtype=T).element_size()
        offs, off = {}, 0
        for name in _STATE_COLS:
            if name in cols:
                b = 8 if name in ("x", "y") else esz
                offs[name] = (off, b)
                off += (n * b + 255) // 256 * 256
        self.state_slab = torch.zeros(max(off, 256), dtype=torch.uint8, device=dev)
        self.state_layout = {name: (o, n, f64 if name in ("x", "y") else T) for name, (o, b) in offs.items()}
        for name in _STATE_COLS:
            if name in offs:
                o, b = offs[name]
                view = self.state_slab[o:o + n * b].view(f64 if name in ("x", "y") else T)
                view.copy_(cols[name])
                setattr(self, name, view)
            else:
                setattr(self, name, None)

    # ---- churn on the device (reference intersection.py:458-539, :576-634) -------------------------
    @classmethod
    def _from_fields(cls, like, n, cols, fields, destq_host, dest_len_host, q_cap):
        g = object.__new__(cls)
        g.model, g.params, g.dtype, g.device = like.model, like.params, like.dtype, like.device
        g.n, g.q_cap = int(n), int(q_cap)
        g._make_state_slab(g.n, cols)
        for name in _FIELD_AXIS:
            setattr(g, name, fields.get(name))
        g.destq_host, g.dest_len_host = destq_host, dest_len_host
        g.hist_cap = like.hist_cap
        g._init_status()
        g.payload_offset = 0
        g._cstate = g._cparams = None
        return g

    @staticmethod
    def _regroup(parts, n_new, q_cap):
        """New group made of ``parts`` = [(group, device index tensor or None, count, first row in the new
        group)]: every per-agent array -- state columns, navigation machine, history rings, dynamic state,
        destination queues, stochastic-rider state -- is re-laid by the library's gather kernel
        (``csf_gather_segments``: one launch per part), not by tensor indexing."""
        like = parts[0][0]
        dev = like.device
        cols = {name: torch.zeros(n_new, dtype=getattr(like, name).dtype, device=dev) for name in _STATE_COLS
                if getattr(like, name) is not None}
        fields = {}
        for name, ax in _FIELD_AXIS.items():
            t = getattr(like, name)
            if t is None:
                continue
            shape = list(t.shape)
            shape[ax] = n_new
            if name == "destq":
                shape[1] = q_cap
            fields[name] = torch.zeros(shape, dtype=t.dtype, device=dev)
        new = AgentGroup._from_fields(like, n_new, cols, fields,
                                      np.zeros((n_new, q_cap, 3)), np.zeros(n_new, dtype=np.int32), q_cap)
        lib = _lib.load()
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            for g, idx, count, off in parts:
                if count == 0:
                    continue
                segs = _lib.CsfGatherSegments()
                k = 0

                def add(src, dst, outer, inner, dst_inner=0):
                    nonlocal k
                    sg = segs.seg[k]
                    sg.src, sg.dst = src.data_ptr(), dst.data_ptr()
                    sg.idx = idx.data_ptr() if idx is not None else None
                    sg.outer, sg.n_src, sg.n_dst, sg.dst_off, sg.count = outer, g.n, n_new, off, count
                    sg.inner_bytes, sg.dst_inner_bytes = inner, dst_inner
                    k += 1

                for name in _STATE_COLS:
                    if getattr(g, name) is not None:
                        add(getattr(g, name), getattr(new, name), 1, getattr(g, name).element_size())
                for name, ax in _FIELD_AXIS.items():
                    t = getattr(g, name)
                    if t is None:
                        continue
                    d = getattr(new, name)
                    if ax == 0:
                        inner = t.element_size() * (t.numel() // max(t.shape[0], 1)) if t.dim() > 1 else t.element_size()
                        dinner = d.element_size() * (d.numel() // max(d.shape[0], 1)) if d.dim() > 1 else d.element_size()
                        add(t, d, 1, inner, 0 if dinner == inner else dinner)
                    else:
                        add(t, d, t.shape[0], t.element_size())
                segs.n = k
                _lib.check(lib.csf_gather_segments(C.byref(segs), st), "csf_gather_segments")
        return new

    def select(self, keep):
        """New group with the agents ``keep`` (indices, any order) -- every per-agent field, including
        the navigation machine, the history rings and the dynamic state, gathered on the device."""
        keep_h = np.asarray(keep, dtype=np.int64).reshape(-1)
        idx = torch.as_tensor(keep_h, device=self.device)
        new = AgentGroup._regroup([(self, idx, len(keep_h), 0)], len(keep_h), self.q_cap)
        new.destq_host, new.dest_len_host = self.destq_host[keep_h].copy(), self.dest_len_host[keep_h].copy()
        return new

    @staticmethod
    def concat(a, b):
        """Agents of ``a`` followed by those of ``b`` (same model and parameter set), joined on the device."""
        assert a.model == b.model and a.dtype == b.dtype
        q_cap = max(a.q_cap, b.q_cap)
        new = AgentGroup._regroup([(a, None, a.n, 0), (b, None, b.n, a.n)], a.n + b.n, q_cap)

        def padq(g):
            hq = np.zeros((g.n, q_cap, 3))
            hq[:, :g.q_cap] = g.destq_host
            return hq

        new.destq_host = np.concatenate([padq(a), padq(b)])
        new.dest_len_host = np.concatenate([a.dest_len_host, b.dest_len_host])
        return new

    # ---- C structs ------------------------------------------------------------------------
    def cstate(self):
        if self._cstate is None:
            s = _lib.CsfAgentState()
            s.n = self.n
            s.first, s.count = 0, self.n
            s.payload_offset = self.payload_offset
            for name in ("x", "y", "psi", "v", "delta", "theta", "deltadot", "thetadot", "vd_default",
                         "step_i", "destq", "dest_len", "dest_ptr", "znav", "znav_v0", "znav_d0",
                         "znav_d1", "prev_x", "prev_y", "hist_x", "hist_y", "hist_step", "ip_x",
                         "ip_zrid", "ip_delta_run", "dyn_x", "dyn_v", "br_gains", "br_poles", "br_vlast", "br_draws",
                         "br_stream", "status"):
                t = getattr(self, name)
                setattr(s, name, t.data_ptr() if t is not None else None)
            s.status_host = self.status_host.data_ptr()     # (UVA: pinned host memory is device-addressable)
            self._cstate = s
        return self._cstate

    def cparams(self, q_scale, origin=(0.0, 0.0)):
        c = self._cparams
        if (c is None or c.q_scale != q_scale or c.q_cap != self.q_cap or c.q_origin[0] != origin[0]
                or c.q_origin[1] != origin[1]):
            self._cparams = self.params.to_agent_params(q_scale, self.q_cap, self.hist_cap, origin)
        return self._cparams

    def invalidate(self):
        self._cstate = None
        self._cparams = None
        self.version += 1

    # ---- host views -----------------------------------------------------------------------
    def states_numpy(self):
        """(n, N_STATES) float64 array in the reference's state order."""
        cols = [self.x, self.y, self.psi, self.v]
        if self.model in ("twod", "invpendulum", "bicycle", "balancingrider"):
            cols.append(self.delta)
        if self.model in ("invpendulum", "balancingrider"):
            cols.append(self.theta)
        if self.model == "balancingrider":
            cols += [self.deltadot, self.thetadot]
        return torch.stack([c.to(torch.float64) for c in cols], dim=1).cpu().numpy()

    def set_state_row(self, k, s):
        """Overwrite the CSF state of agent k (host -> device)."""
        s = np.asarray(s, float)
        self.x[k], self.y[k] = float(s[0]), float(s[1])
        self.psi[k], self.v[k] = float(_wrap_angle(s[2])), float(s[3])
        for name, idx in (("delta", 4), ("theta", 5), ("deltadot", 6), ("thetadot", 7)):
            t = getattr(self, name)
            if t is not None and len(s) > idx:
                t[k] = float(s[idx])

    def set_destqueue(self, k, q):
        q = np.asarray(q, float).reshape(-1, 3)
        if len(q) > self.q_cap:
            new_cap = max(len(q), 2 * self.q_cap)
            host = np.zeros((self.n, new_cap, 3))
            host[:, :self.q_cap] = self.destq_host
            self.destq_host = host
            self.q_cap = new_cap
            self.destq = torch.as_tensor(host, dtype=torch.float64, device=self.device)
            self.invalidate()
        self.destq_host[k] = 0
        self.destq_host[k, :len(q)] = q
        self.dest_len_host[k] = len(q)
        self.destq[k] = torch.as_tensor(self.destq_host[k], dtype=torch.float64, device=self.device)
        self.dest_len[k] = len(q)

    _RECORD_FIELDS = ("x", "y", "psi", "v", "delta", "theta", "deltadot", "thetadot", "step_i", "dest_ptr",
                      "znav", "znav_v0", "znav_d0", "znav_d1", "prev_x", "prev_y", "hist_x", "hist_y",
                      "hist_step", "ip_x", "ip_zrid", "ip_delta_run", "dyn_x", "dyn_v", "br_gains", "br_poles", "br_vlast",
                      "br_draws", "br_stream")

    def export_records(self):
        """Full per-agent device state as host dicts (kept by vehicles across re-binding)."""
        host = {n: getattr(self, n).cpu().numpy() for n in self._RECORD_FIELDS if getattr(self, n) is not None}
        states = self.states_numpy()
        recs = []
        for k in range(self.n):
            r = {"_s": states[k].copy(), "model": self.model}
            for n, a in host.items():
                r[n] = a[..., k].copy() if a.ndim == 2 else a[k].copy()
            recs.append(r)
        return recs

    def import_record(self, k, rec):
        if rec.get("model") != self.model:
            return
        for n in self._RECORD_FIELDS:
            t = getattr(self, n)
            if t is None or n not in rec:
                continue
            val = torch.as_tensor(np.asarray(rec[n]), dtype=t.dtype, device=t.device)
            if t.dim() == 2:
                t[:, k] = val
            else:
                t[k] = val

    def cstate_range(self, first, count):
        s = _lib.CsfAgentState()
        C.memmove(C.byref(s), C.byref(self.cstate()), C.sizeof(s))
        s.first, s.count = first, count
        return s

    def poll_status(self):
        """True if a kernel has raised a status bit since the last check (no copy, no synchronisation)."""
        return bool(self.status_host[0] != 0)

    def check_status(self):
        st = int(self.status.item())
        if st & 1:
            raise FloatingPointError("non-finite force or state on the device (status bit 0)")
        if st & 2:
            raise RuntimeError("Invalid navigation state")             # vehicle.py:455
        if st & 4:
            raise OverflowError("agent position left the Q-format range of the f32 payload")
        return st


class ObstacleGroup:
    """Road users that exert a force but are not stepped by the kernels
    (UncontrolledVehicle, reference vehicle.py:920-987).  Host-owned x, y, psi."""

    model = "uncontrolled"

    def __init__(self, xypsi, params=None, device="cuda"):
        self.params = params if params is not None else VehicleParameters()
        self.device = torch.device(device)
        self.set(xypsi)
        self.payload_offset = 0

    def set(self, xypsi):
        a = np.atleast_2d(np.asarray(xypsi, float))
        self.n = a.shape[0]
        self.host = a[:, :3].copy()
        self.x = torch.as_tensor(a[:, 0].copy(), dtype=torch.float64, device=self.device)
        self.y = torch.as_tensor(a[:, 1].copy(), dtype=torch.float64, device=self.device)
        self.psi = torch.as_tensor(_wrap_angle(a[:, 2]), dtype=torch.float64, device=self.device)


class Engine:
    """One interaction domain (``SocialForceIntersection``): groups + obstacles + road edges."""

    def __init__(self, groups, obstacles=None, priority_rule="unregulated", road_edges=(),
                 dtype=torch.float32, device="cuda", q_scale=None, extent=None, origin=None, scenario_size=None,
                 n_global=None, global_offset=0, exchange=None, pair_mode="auto", resort_every=64,
                 count_pairs=False, graph=False, global_classes=None, reuse=None):
        """``q_scale`` / ``extent`` / ``origin``: the Q-format frame of the f32 payload -- positions are
        stored as int32 multiples of ``q_scale`` (default: the finest power of two that fits ``extent``
        into 2^30 units) relative to ``origin`` (default with ``extent``/``q_scale``: (0, 0); with neither:
        frame fitted to the crowd, see ``parameters.payload_frame``).
        ``n_global`` / ``global_offset`` / ``exchange``: agent-range sharding of one crowd over
        several GPUs -- this engine owns agents [global_offset, global_offset + n) of an
        ``n_global``-agent crowd; ``exchange(payload)`` is called after every step to all-gather the pair
        payload (see distributed.py).  The field parameters of the whole crowd are those of this engine's
        first group unless ``global_classes`` = [(first, count, params), ...] names the parameter set of
        every contiguous range of the GLOBAL numbering (the same list on every rank: a source's field
        parameters are not part of the exchanged payload, so every rank has to know them up front).
        ``reuse``: an engine that is being replaced (road-user churn): its device buffers -- payload, force
        arrays, sorted copy, tiles, permutations, workspace -- are taken over wherever they are large enough
        (buffers are allocated with 25 % head room), so that adding or removing road users does not go through
        the allocator.
        ``graph=True``: ``step()`` replays a CUDA graph of the step's kernel sequence (captured at the
        first call; the periodic re-sort of the spatial order and the exchange stay outside it)."""
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.CsfError("no CUDA device: the csf_b200 engine has no CPU fallback")
        self.device = torch.device(device)
        self._pool = {}
        if reuse is not None and reuse.device == self.device and reuse.dtype == dtype:
            self._pool = reuse._pool
            reuse._pool = {}
            reuse._graph = None
        self.dtype = dtype
        self.f32 = dtype == torch.float32
        self.sfx = "f32" if self.f32 else "f64"
        self.groups = [g for g in groups if g.n > 0]
        self.obstacles = [o for o in (obstacles or []) if o.n > 0]
        self.p2r = priority_rule == "p2r"
        self.scenario_size = scenario_size
        self.exchange = exchange
        # pair kernel choice: "dense" (thread-per-target, every pair), "tiled" (spatially tiled
        # sources + field-of-view culling), "auto" = tiled from 2048 road users on
        assert pair_mode in ("auto", "dense", "tiled")
        self.pair_mode = pair_mode
        self.resort_every = int(resort_every)
        self.use_graph = bool(graph)
        self._graph = None
        self.global_offset = int(global_offset)
        off = self.global_offset
        for g in self.groups + self.obstacles:
            g.payload_offset = off
            if hasattr(g, "invalidate"):
                g.invalidate()
            off += g.n
        self.n_agents = sum(g.n for g in self.groups)
        self.n_total = off if n_global is None else int(n_global)
        self._sharded = n_global is not None
        # Q-format scale of the f32 payload
        if q_scale is None and extent is None and n_global is not None:
            # every rank must quantise the exchanged payload with the same scale: it cannot be derived
            # from the local shard
            raise ValueError("a sharded crowd (n_global) needs q_scale or extent: the same value on every rank")
        if q_scale is None:
            if extent is None:
                pts = []
                for g in self.groups:
                    pts.append(torch.stack([g.x, g.y], dim=1).cpu().numpy())
                    valid = np.arange(g.q_cap)[None, :] < g.dest_len_host[:, None]
                    pts.append(g.destq_host[..., :2][valid])
                pts += [o.host[:, :2] for o in self.obstacles]
                fitted, extent = payload_frame(pts)
                if origin is None:
                    origin = fitted
            q_scale = choose_q_scale(extent)
        self.q_scale = float(q_scale)
        self.q_origin = (0.0, 0.0) if origin is None else (float(origin[0]), float(origin[1]))
        self.elem_bytes = 16 if self.f32 else 32
        n = max(self.n_total, 1)
        if exchange is not None and hasattr(exchange, "payload_tensor"):
            self.payload = exchange.payload_tensor()      # lives in the peer-shared buffer (PeerExchange)
            assert self.payload.shape[0] == n and self.payload.dtype == (torch.int32 if self.f32 else torch.float64)
        else:
            self.payload = self._buf("payload", (n, 4), torch.int32 if self.f32 else torch.float64)
        self.frep = self._buf("frep", (max(self.n_agents, 1), 2), dtype)
        self.force = self._buf("force", (max(self.n_agents, 1), 2), dtype)
        self.fdest = self._buf("fdest", (max(self.n_agents, 1), 2), dtype)
        self.froad = None
        self.set_road_edges(road_edges)
        # source classes: contiguous payload ranges with identical field parameters
        self.classes = []
        if n_global is not None:
            if any(g.model == "bicycle" for g in self.groups):
                raise NotImplementedError("v0.1 Bicycle-field sources in a sharded crowd: their eccentricity "
                                          "depends on the speed, which is not part of the exchanged payload")
            if global_classes is None:
                global_classes = [(0, self.n_total, self.groups[0].params)]
            covered = 0
            for first, count, par in global_classes:
                if int(first) != covered or count <= 0:
                    raise ValueError("global_classes must tile [0, n_global) with contiguous ranges")
                covered += int(count)
                key = par.field_key()
                if self.classes and self.classes[-1][2] == key:
                    s0_, c0_, k0_, fp0_ = self.classes[-1]
                    self.classes[-1] = (s0_, c0_ + int(count), k0_, fp0_)
                else:
                    self.classes.append((int(first), int(count), key, par.to_field_params(self.q_scale, self.p2r)))
            if covered != self.n_total:
                raise ValueError("global_classes must tile [0, n_global) with contiguous ranges")
        for g in (self.groups + self.obstacles if n_global is None else []):
            kind = 1 if g.model == "bicycle" else 0      # v0.1 elliptic field (vehicle.py:1107-1147)
            key = g.params.field_key(kind)
            if self.classes and self.classes[-1][2] == key and kind == 0:
                s, c, k, fp = self.classes[-1]
                self.classes[-1] = (s, c + g.n, k, fp)
            else:
                self.classes.append((g.payload_offset, g.n, key,
                                     g.params.to_field_params(self.q_scale, self.p2r, kind)))
        # per-source eccentricity of Bicycle-field sources (refreshed every step from their speed)
        self._ecc = {}
        for g in self.groups:
            if g.model == "bicycle":
                self._ecc[g.payload_offset] = (g, torch.zeros(g.n, dtype=dtype, device=self.device))
        wsb = 0
        if self.n_total > 1 and scenario_size is None:
            for s, c, _, _ in self.classes:
                wsb = max(wsb, int(self.lib.csf_pair_workspace_bytes(c, self.n_agents, 4 if self.f32 else 8)))
        # tiled + culled kernel for every source class, TwoD field or v0.1 Bicycle field (the latter's sorted
        # copy carries the headings scaled by the eccentricities: csf_tile_sources_bicycle)
        self.tiled = (scenario_size is None and self.n_total > 1 and
                      (pair_mode == "tiled" or (pair_mode == "auto" and self.n_total >= 2048)))
        self.pair_stats = torch.zeros(16 + 4 * 16384, dtype=torch.int64, device=self.device) if count_pairs else None
        self._tiles = []
        self._pair_calls = 0
        if self.tiled:
            eb = 4 if self.f32 else 8
            tile_elems = self.lib.csf_tiled_tile_bytes(eb) // eb
            for ci, (s, c, _, _) in enumerate(self.classes):
                n_pad = int(self.lib.csf_tiled_padded_sources(c))
                n_tiles = int(self.lib.csf_tiled_num_tiles(c))
                self._tiles.append(dict(
                    sorted=self._buf(f"sorted{ci}", (n_pad, 4), self.payload.dtype),
                    tiles=self._buf(f"tiles{ci}", (n_tiles, tile_elems), self.payload.dtype),
                    perm=self._buf(f"perm{ci}", c, torch.int64)))
                # scheduling of the pair kernel's work items: cost per item (written by every launch) and
                # the order to hand them out in (heaviest first; refreshed with the spatial order)
                n_items = int(self.lib.csf_tiled_num_items(c, self.n_agents, eb))
                tl = self._tiles[-1]
                tl["n_items"] = n_items
                tl["item_cost"] = self._buf(f"item_cost{ci}", max(n_items, 1), torch.int32)
                tl["item_order"] = (torch.arange(n_items, dtype=torch.int32, device=self.device)
                                    if 0 < n_items <= 4096 else None)
                wsb = max(wsb, int(self.lib.csf_pair_tiled_workspace_bytes(c, self.n_agents, eb)))
            # visiting order of the local targets (Morton order too: compact target blocks)
            self._key_box = None
            self._tgt_perm = None
            self._order_valid = False
            self._single_class = (len(self.classes) == 1 and self.classes[0][0] == self.global_offset
                                  and self.classes[0][1] == self.n_agents)
            self._morton = (-float(extent if extent is not None else 2.0 ** 30 * self.q_scale),
                            2.0 * float(extent if extent is not None else 2.0 ** 30 * self.q_scale) / 65536.0)
        self.ws = self._buf("ws", max(wsb, 16), torch.uint8, zero=False)
        # Fused step (three launches: tile build + block bounds [+ wait for the peers' pushes] -> pair kernel
        # -> per-agent kernel that also reduces the pair kernel's partial sums [and signals / pushes to the
        # peers]): one tiled source class, no Bicycle-field sources
        self._fused = bool(self.tiled and len(self.classes) == 1 and self.classes[0][3].field_kind == 0)
        self._fusion = {}
        self.gpu_launches = 0
        self._road_version = 0
        self._graph_version = None
        with torch.cuda.device(self.device):
            self.pack()

    # ---- helpers ----------------------------------------------------------------------------
    def _buf(self, name, shape, dtype, zero=True):
        """Device array ``shape`` of ``dtype``, carved from the pooled buffer ``name`` if that is large enough
        (see ``reuse``); otherwise a new buffer with 25 % head room replaces it."""
        shape = tuple(int(d) for d in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        numel = 1
        for d in shape:
            numel *= d
        flat = self._pool.get(name)
        if flat is None or flat.dtype != dtype or flat.numel() < numel:
            flat = torch.empty(max(numel + numel // 4, 16), dtype=dtype, device=self.device)
            self._pool[name] = flat
        out = flat[:numel].view(shape)
        if zero:
            out.zero_()
        return out

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _fn(self, name):
        return getattr(self.lib, f"{name}_{self.sfx}")

    def set_road_edges(self, road_edges):
        """road_edges: iterable of (vertices (M,2), F_0, sigma)."""
        merged = {}
        for verts, F_0, sigma in road_edges:
            merged.setdefault((float(F_0), float(sigma)), []).append(np.asarray(verts, float).reshape(-1, 2))
        self.road = [(torch.as_tensor(np.ascontiguousarray(np.concatenate(v)), dtype=torch.float64,
                                      device=self.device).contiguous(), k[0], k[1])
                     for k, v in merged.items()]
        self.froad = self._buf("froad", tuple(self.frep.shape), self.frep.dtype) if self.road else None
        self._road_version = getattr(self, "_road_version", 0) + 1

    def _version(self):
        """Changes whenever something a captured step graph has baked in (device pointers of the groups,
        their parameters, the road vertices) is replaced."""
        return (self._road_version,) + tuple(g.version for g in self.groups)

    def pack(self):
        """update_road_user_positions (intersection.py:660-677): state -> pair payload."""
        st = self._stream()
        for g in self.groups:
            _lib.check(self._fn("csf_pack_xycs")(C.byref(g.cstate()), C.byref(g.cparams(self.q_scale, self.q_origin)),
                                                 _ptr(self.payload), st), "csf_pack_xycs")
            self.gpu_launches += 1
        for o in self.obstacles:
            dst = C.c_void_p(self.payload.data_ptr() + o.payload_offset * self.elem_bytes)
            _lib.check(self._fn("csf_pack_xypsi")(_ptr(o.x), _ptr(o.y), _ptr(o.psi), o.n, self.q_scale, self.q_origin[0],
                                                  self.q_origin[1], dst, st),
                       "csf_pack_xypsi")
            self.gpu_launches += 1

    _capturing = False

    def _refresh_order(self):
        """Spatial (Hilbert) visiting order of the sources of every class and of the local targets, on the
        device and inside the library: bounding box of all road users, keys, radix sort (csf_spatial_*),
        written into buffers whose addresses never change (a captured CUDA graph keeps pointing at them)."""
        st = self._stream()
        if self._key_box is None:
            self._key_box = self._buf("key_box", 4, torch.float64)
            nmax = max([c for _, c, _, _ in self.classes] + [self.n_agents])
            self._order_ws = self._buf("order_ws", int(self.lib.csf_spatial_order_workspace_bytes(nmax)), torch.uint8,
                                       zero=False)
        _lib.check(self._fn("csf_spatial_bbox")(_ptr(self.payload), self.n_total, _ptr(self._key_box), st),
                   "csf_spatial_bbox")
        self.gpu_launches += 1
        for ci, (s, c, _, fp) in enumerate(self.classes):
            tl = self._tiles[ci]
            src = C.c_void_p(self.payload.data_ptr() + s * self.elem_bytes)
            _lib.check(self._fn("csf_spatial_order")(src, c, _ptr(self._key_box), _ptr(tl["perm"]), _ptr(self._order_ws),
                                                     self._order_ws.numel(), st), "csf_spatial_order")
            self.gpu_launches += 2
        if self._single_class:      # one class covering exactly the targets: same order
            self._tgt_perm = self._tiles[0]["perm"]
        elif self._sharded and _SHARD_TARGET_ORDER == "partition":
            # A rank's targets are a contiguous range of a numbering along a space-filling curve (the partition,
            # fixed for the run): they are visited in that order.  Sorting them along the curve of the CURRENT
            # bounding box instead looks equivalent but is not -- the range is a segment of the numbering's curve,
            # and along any other curve (the box has moved by a few metres since) it is left and re-entered many
            # times, so that a few blocks of 64 consecutive targets straddle a jump across the whole region: one
            # such block streams every source through its filter, and that one item then IS the launch (measured
            # on a half crowd: pair kernel 0.17 -> 0.27 ms from the first re-sort on).
            if self._tgt_perm is None:
                self._tgt_perm = self._buf("tgt_perm", self.n_agents, torch.int64)
                self._tgt_perm.copy_(torch.arange(self.n_agents, dtype=torch.int64, device=self.device))
        else:
            tgt = C.c_void_p(self.payload.data_ptr() + self.global_offset * self.elem_bytes)
            if self._tgt_perm is None:
                self._tgt_perm = self._buf("tgt_perm", self.n_agents, torch.int64)
            _lib.check(self._fn("csf_spatial_order")(tgt, self.n_agents, _ptr(self._key_box), _ptr(self._tgt_perm),
                                                     _ptr(self._order_ws), self._order_ws.numel(), st),
                       "csf_spatial_order")
            self.gpu_launches += 2
        self._order_valid = True
        self._refresh_item_order()

    def _refresh_item_order(self):
        """Hand the pair kernel's work items out heaviest first (costs measured by the previous launch)."""
        st = self._stream()
        for tl in self._tiles:
            if tl.get("item_order") is not None and _ITEM_ORDER_MODE != "none":
                _lib.check(self.lib.csf_tiled_item_order(_ptr(tl["item_cost"]), tl["n_items"],
                                                         _ptr(tl["item_order"]), st), "csf_tiled_item_order")
                self.gpu_launches += 1

    def _maybe_refresh(self):
        """Outside the step's kernel sequence (and outside its CUDA graph): the spatial order every
        ``resort_every`` steps; the item order also right after the first launch has measured the costs."""
        if not self.tiled:
            return
        if not self._order_valid or self._pair_calls % max(self.resort_every, 1) == 0:
            if self.exchange is not None and hasattr(self.exchange, "begin_step"):
                self.exchange.begin_step()  # the keys are computed from the peers' payload: wait for their pushes
            self._refresh_order()
            self._item_order_at = self._pair_calls
        elif self._pair_calls == (1 if _ITEM_ORDER_MODE == "stale" else getattr(self, "_item_order_at", 0) + 1):
            # the launch that followed the re-sort has measured the items' costs in the new order
            self._refresh_item_order()

    def _pair_and_road(self):
        st = self._stream()
        if self.exchange is not None and hasattr(self.exchange, "begin_step"):
            self.exchange.begin_step()                    # wait for every peer's payload push
        have_rep = self._pair(st)
        if self.exchange is not None and hasattr(self.exchange, "after_pair"):
            self.exchange.after_pair()                    # the payload has been read: peers may overwrite it
        self._road(st)
        return have_rep

    def _pair(self, st):
        have_rep = self.n_total > 1
        if have_rep:
            if self.scenario_size is not None:
                fp = self.classes[0][3]
                _lib.check(self._fn("csf_pair_forces_grouped")(_ptr(self.payload), self.n_agents,
                                                               int(self.scenario_size), C.byref(fp),
                                                               _ptr(self.frep), st), "csf_pair_forces_grouped")
                self.gpu_launches += 1
            else:
                if not self._capturing:
                    self._maybe_refresh()
                self._pair_calls += 1
                tgt = C.c_void_p(self.payload.data_ptr() + self.global_offset * self.elem_bytes)
                for ci, (s, c, _, fp) in enumerate(self.classes):
                    src = C.c_void_p(self.payload.data_ptr() + s * self.elem_bytes)
                    if fp.field_kind == 1 and self.tiled:
                        g, _ = self._ecc[s]
                        tl = self._tiles[ci]
                        _lib.check(self._fn("csf_tile_sources_bicycle")(src, _ptr(g.v), fp.v_max, c, _ptr(tl["perm"]),
                                                                        _ptr(tl["sorted"]), _ptr(tl["tiles"]), st),
                                   "csf_tile_sources_bicycle")
                        _lib.check(self._fn("csf_pair_forces_tiled")(
                            _ptr(tl["sorted"]), _ptr(tl["tiles"]), c, tgt, _ptr(self._tgt_perm), self.n_agents,
                            C.byref(fp), _ptr(self.frep), 1 if ci > 0 else 0, _ptr(self.ws), self.ws.numel(),
                            _ptr(tl["item_order"]), _ptr(tl["item_cost"]), _ptr(self.pair_stats), 0, st),
                            "csf_pair_forces_tiled")
                        self.gpu_launches += 4   # tile build, block bounds, pair, reduce
                        continue
                    if fp.field_kind == 1:
                        g, ecc = self._ecc[s]
                        _lib.check(self._fn("csf_bicycle_eccentricity")(_ptr(g.v), g.n, fp.v_max, _ptr(ecc), st),
                                   "csf_bicycle_eccentricity")
                        _lib.check(self._fn("csf_pair_forces_bicycle")(src, _ptr(ecc), c, tgt, self.n_agents,
                                                                       C.byref(fp), _ptr(self.frep),
                                                                       1 if ci > 0 else 0, st),
                                   "csf_pair_forces_bicycle")
                        self.gpu_launches += 2
                        continue
                    if self.tiled:
                        tl = self._tiles[ci]
                        _lib.check(self._fn("csf_tile_sources")(src, c, _ptr(tl["perm"]), _ptr(tl["sorted"]),
                                                                _ptr(tl["tiles"]), st), "csf_tile_sources")
                        _lib.check(self._fn("csf_pair_forces_tiled")(
                            _ptr(tl["sorted"]), _ptr(tl["tiles"]), c, tgt, _ptr(self._tgt_perm), self.n_agents,
                            C.byref(fp), _ptr(self.frep), 1 if ci > 0 else 0, _ptr(self.ws), self.ws.numel(),
                            _ptr(tl["item_order"]), _ptr(tl["item_cost"]), _ptr(self.pair_stats), 0, st),
                            "csf_pair_forces_tiled")
                        self.gpu_launches += 4   # tile build (+ chunk bounds), block bounds, pair, reduce
                        continue
                    _lib.check(self._fn("csf_pair_forces")(src, c, tgt, self.n_agents,
                                                           C.byref(fp), _ptr(self.frep), 1 if ci > 0 else 0,
                                                           _ptr(self.ws), self.ws.numel(), st), "csf_pair_forces")
                    self.gpu_launches += 2
        return have_rep

    def _road(self, st):
        if self.road:
            for g in self.groups:
                out = self._off(self.froad, g)
                for ei, (verts, F_0, sigma) in enumerate(self.road):
                    _lib.check(self._fn("csf_road_forces")(_ptr(g.x), _ptr(g.y), g.n, _ptr(verts), verts.shape[0],
                                                           F_0, sigma, out, 1 if ei > 0 else 0, st),
                               "csf_road_forces")
                    self.gpu_launches += 1

    def _fusion_of(self, g):
        """CsfStepFusion of group g: where the pair kernel leaves its partial sums, and the peer exchange."""
        key = (id(g), g.payload_offset)
        fu = self._fusion.get(key)
        if fu is None:
            s, c, _, fp = self.classes[0]
            eb = 4 if self.f32 else 8
            fu = _lib.CsfStepFusion()
            fu.partial = self.ws.data_ptr() + int(self.lib.csf_tiled_partial_offset(c, self.n_agents, eb))
            fu.partial_stride = self.n_agents
            fu.partial_offset = g.payload_offset - self.global_offset
            fu.n_groups = int(self.lib.csf_tiled_num_groups(c, self.n_agents, eb))
            fu.f0 = fp.f_0
            # the per-agent kernel follows the pair kernel directly in the stream: launched as its programmatic
            # dependent, its destination-force part runs in the pair kernel's tail (CSF_PDL=0: plain launch)
            fu.pdl = 0 if os.environ.get("CSF_PDL", "1") == "0" else 1
            if self._peer_fused():
                fu.comm = self.exchange.comm
            self._fusion[key] = fu
        return fu

    def _peer_fused(self):
        return self.exchange is not None and bool(getattr(self.exchange, "fused", False))

    def _step_fused(self, exchange=True, mark=None):
        """The step as three launches (see __init__).  ``mark(i)``, if given, is called before launch i and
        after the last one (bench.py records CUDA events there: per-kernel durations)."""
        mark = mark or (lambda i: None)
        st = self._stream()
        if not self._capturing:
            self._maybe_refresh()
        self._pair_calls += 1
        s, c, _, fp = self.classes[0]
        tl = self._tiles[0]
        src = C.c_void_p(self.payload.data_ptr() + s * self.elem_bytes)
        tgt = C.c_void_p(self.payload.data_ptr() + self.global_offset * self.elem_bytes)
        comm = C.byref(self.exchange.comm) if self._peer_fused() else None
        if self.exchange is not None and not self._peer_fused() and hasattr(self.exchange, "begin_step"):
            self.exchange.begin_step()
        mark(0)
        _lib.check(self._fn("csf_tiled_prepare")(src, c, _ptr(tl["perm"]), _ptr(tl["sorted"]), _ptr(tl["tiles"]), tgt,
                                                 _ptr(self._tgt_perm), self.n_agents, _ptr(self.ws), self.ws.numel(),
                                                 comm, st), "csf_tiled_prepare")
        mark(1)
        _lib.check(self._fn("csf_pair_forces_tiled")(
            _ptr(tl["sorted"]), _ptr(tl["tiles"]), c, tgt, _ptr(self._tgt_perm), self.n_agents, C.byref(fp),
            _ptr(self.frep), 0, _ptr(self.ws), self.ws.numel(), _ptr(tl["item_order"]), _ptr(tl["item_cost"]),
            _ptr(self.pair_stats), _lib.CSF_TILED_PREPARED | _lib.CSF_TILED_NO_REDUCE |
            (0 if os.environ.get("CSF_PDL", "1") == "0" else _lib.CSF_TILED_PDL), st), "csf_pair_forces_tiled")
        self.gpu_launches += 2
        mark(2)
        if self.exchange is not None and not self._peer_fused() and hasattr(self.exchange, "after_pair"):
            self.exchange.after_pair()
        self._road(st)
        for g in self.groups:
            _lib.check(self._fn("csf_agent_step_fused")(
                _lib.MODEL_IDS[g.model], C.byref(g.cstate()), C.byref(g.cparams(self.q_scale, self.q_origin)),
                self.n_total, C.byref(self._fusion_of(g)), self._off(self.froad, g), self._off(self.force, g),
                _ptr(self.payload), st), "csf_agent_step_fused")
            self.gpu_launches += 1
        mark(3)
        if exchange and self.exchange is not None and not self._peer_fused():
            self.exchange(self.payload)

    def _pair_alone(self, mark=None):
        """Measurement aid (bench.py on a sharded crowd): the tile build WITHOUT the wait for the peers, then the
        pair kernel, on whatever the payload buffer holds -- no reduction, no per-agent kernel, no exchange; the
        partial sums it leaves are overwritten by the next step.  Call it only while every rank is quiescent."""
        assert self._fused and self._order_valid
        mark = mark or (lambda i: None)
        st = self._stream()
        s, c, _, fp = self.classes[0]
        tl = self._tiles[0]
        src = C.c_void_p(self.payload.data_ptr() + s * self.elem_bytes)
        tgt = C.c_void_p(self.payload.data_ptr() + self.global_offset * self.elem_bytes)
        mark(0)
        _lib.check(self._fn("csf_tiled_prepare")(src, c, _ptr(tl["perm"]), _ptr(tl["sorted"]), _ptr(tl["tiles"]), tgt,
                                                 _ptr(self._tgt_perm), self.n_agents, _ptr(self.ws), self.ws.numel(),
                                                 None, st), "csf_tiled_prepare")
        mark(1)
        _lib.check(self._fn("csf_pair_forces_tiled")(
            _ptr(tl["sorted"]), _ptr(tl["tiles"]), c, tgt, _ptr(self._tgt_perm), self.n_agents, C.byref(fp),
            _ptr(self.frep), 0, _ptr(self.ws), self.ws.numel(), _ptr(tl["item_order"]), _ptr(tl["item_cost"]),
            _ptr(self.pair_stats), _lib.CSF_TILED_PREPARED | _lib.CSF_TILED_NO_REDUCE, st), "csf_pair_forces_tiled")
        mark(2)
        self.gpu_launches += 2

    def _step_kernels(self, exchange=True):
        if self._fused and self.n_total > 1:
            self._step_fused(exchange)
        else:
            self._agent_step(self._pair_and_road(), exchange=exchange)

    def _off(self, t, g):
        if t is None:
            return C.c_void_p(0)
        return C.c_void_p(t.data_ptr() + (g.payload_offset - self.global_offset) * 2 * t.element_size())

    # ---- the three public operations -----------------------------------------------------------
    def calc_forces(self):
        """calc_forces (intersection.py:747-864): returns the device tensor force (N, 2)."""
        if self.poll_status():
            self.check_status()
        have_rep = self._pair_and_road()
        st = self._stream()
        for g in self.groups:
            _lib.check(self._fn("csf_agent_forces")(
                _lib.MODEL_IDS[g.model], C.byref(g.cstate()), C.byref(g.cparams(self.q_scale, self.q_origin)), self.n_total,
                self._off(self.frep, g) if have_rep else C.c_void_p(0), self._off(self.froad, g),
                self._off(self.force, g), self._off(self.fdest, g), st), "csf_agent_forces")
            self.gpu_launches += 1
        return self.force

    def advance(self):
        """vehicles[i].step(Fx[i], Fy[i]) for all i + update_road_user_positions (:891-894)."""
        st = self._stream()
        for g in self.groups:
            _lib.check(self._fn("csf_agent_advance")(
                _lib.MODEL_IDS[g.model], C.byref(g.cstate()), C.byref(g.cparams(self.q_scale, self.q_origin)),
                self._off(self.force, g), _ptr(self.payload), st), "csf_agent_advance")
            self.gpu_launches += 1

    def _agent_step(self, have_rep, exchange=True):
        st = self._stream()
        for g in self.groups:
            _lib.check(self._fn("csf_agent_step")(
                _lib.MODEL_IDS[g.model], C.byref(g.cstate()), C.byref(g.cparams(self.q_scale, self.q_origin)), self.n_total,
                self._off(self.frep, g) if have_rep else C.c_void_p(0), self._off(self.froad, g),
                self._off(self.force, g), _ptr(self.payload), st), "csf_agent_step")
            self.gpu_launches += 1
        if exchange and self.exchange is not None:
            self.exchange(self.payload)

    def step(self):
        """SocialForceIntersection.step (intersection.py:866-896), fused per-agent kernel."""
        if self.n_agents == 0:
            return
        if self.poll_status():
            self.check_status()                 # raises: NaN / navigation state / payload range / exchange time-out
        with torch.cuda.device(self.device):
            if self.use_graph:
                return self._step_graph()
            self._step_kernels()

    def poll_status(self):
        """True if any kernel has raised a status bit (host-mapped words: no copy, no synchronisation)."""
        if any(g.poll_status() for g in self.groups):
            return True
        return bool(self.exchange is not None and hasattr(self.exchange, "poll_status") and self.exchange.poll_status())

    def _step_graph(self):
        """The same kernel sequence replayed from a CUDA graph (launch-bound small crowds / shards)."""
        self._maybe_refresh()
        self._exchange_in_graph = bool(getattr(self.exchange, "capturable", False))
        if self._graph is not None and self._graph_version != self._version():
            self._graph = None              # a destination queue grew / road edges changed: stale device pointers
        if self._graph is None:
            launches0 = self.gpu_launches
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            g = torch.cuda.CUDAGraph()
            self._capturing = True
            try:
                with torch.cuda.stream(side):
                    calls = self._pair_calls
                    with torch.cuda.graph(g, stream=side):
                        self._step_kernels(exchange=self._exchange_in_graph)
                    self._pair_calls = calls
            finally:
                self._capturing = False
            torch.cuda.current_stream(self.device).wait_stream(side)
            self._graph_launches = self.gpu_launches - launches0
            self.gpu_launches = launches0
            self._graph = g
            self._graph_version = self._version()
        self._graph.replay()
        self._pair_calls += 1
        self.gpu_launches += self._graph_launches
        if self.exchange is not None and not self._exchange_in_graph:
            self.exchange(self.payload)

    def step_host(self, host_in, host_out, host_force=None):
        """One step driven from HOST buffers (pinned): upload the CSF state (x, y, psi, v, delta ...)
        of every group-0 agent, step, download the new state and the total force.  This is the
        call a host-side co-simulation loop makes; bench.py's ``e2e`` times it.
        With ``graph=True`` and slab buffers the whole sequence -- upload, payload re-pack (+ exchange),
        the step's kernels, both downloads -- is ONE CUDA graph per buffer set: one launch and one
        synchronisation per step."""
        g = self.groups[0]
        if self.poll_status():
            self.check_status()
        with torch.cuda.device(self.device):
            if self.use_graph and torch.is_tensor(host_in) and torch.is_tensor(host_out):
                return self._step_host_graph(host_in, host_out, host_force)
            self._host_upload(g, host_in)
            self.step()
            self._host_download(g, host_out, host_force)
            torch.cuda.current_stream(self.device).synchronize()

    def _host_upload(self, g, host_in):
        if torch.is_tensor(host_in):                # one pinned slab (layout: g.state_layout): one copy each way
            g.state_slab.copy_(host_in, non_blocking=True)
        else:
            for name, src in host_in.items():
                getattr(g, name).copy_(src, non_blocking=True)
        self.pack()
        if self.exchange is not None:
            self.exchange(self.payload)

    def _host_download(self, g, host_out, host_force):
        if torch.is_tensor(host_out):
            host_out.copy_(g.state_slab, non_blocking=True)
        else:
            for name, dst in host_out.items():
                dst.copy_(getattr(g, name), non_blocking=True)
        if host_force is not None:
            host_force.copy_(self.force[:g.n], non_blocking=True)

    def _step_host_graph(self, host_in, host_out, host_force):
        g = self.groups[0]
        capt = bool(getattr(self.exchange, "capturable", self.exchange is None))
        if not capt:                                 # a collective that cannot be captured: kernel by kernel
            self._host_upload(g, host_in)
            self.step()
            self._host_download(g, host_out, host_force)
            return torch.cuda.current_stream(self.device).synchronize()
        self._maybe_refresh()
        key = (host_in.data_ptr(), host_out.data_ptr(), host_force.data_ptr() if host_force is not None else 0,
               self._version())
        cache = self.__dict__.setdefault("_host_graphs", {})
        entry = cache.get(key)
        if entry is None:
            if len(cache) > 8:
                cache.clear()
            launches0 = self.gpu_launches
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            gr = torch.cuda.CUDAGraph()
            self._capturing = True
            try:
                with torch.cuda.stream(side):
                    calls = self._pair_calls
                    with torch.cuda.graph(gr, stream=side):
                        self._host_upload(g, host_in)
                        self._step_kernels(exchange=True)
                        self._host_download(g, host_out, host_force)
                    self._pair_calls = calls
            finally:
                self._capturing = False
            torch.cuda.current_stream(self.device).wait_stream(side)
            entry = cache[key] = (gr, self.gpu_launches - launches0)
            self.gpu_launches = launches0
        entry[0].replay()
        self._pair_calls += 1
        self.gpu_launches += entry[1]
        torch.cuda.current_stream(self.device).synchronize()

    def check_status(self):
        if self.exchange is not None and hasattr(self.exchange, "check_status"):
            self.exchange.check_status()
        return [g.check_status() for g in self.groups]

    # ---- per-vehicle operations (the reference's per-vehicle hooks) -------------------------------
    def agent_dest_force(self, g, k):
        """Vehicle.calcDestinationForce() of one agent (mutates its navigation state)."""
        st = g.cstate_range(k, 1)
        _lib.check(self._fn("csf_agent_forces")(
            _lib.MODEL_IDS[g.model], C.byref(st), C.byref(g.cparams(self.q_scale, self.q_origin)), 1, C.c_void_p(0),
            C.c_void_p(0), self._off(self.force, g), self._off(self.fdest, g), self._stream()), "csf_agent_forces")
        self.gpu_launches += 1
        f = self.fdest[g.payload_offset - self.global_offset + k].to(torch.float64).cpu().numpy()
        return float(f[0]), float(f[1])

    def agent_advance(self, g, k, F1, F2):
        """Vehicle.step(F1, F2) of one agent."""
        self.force[g.payload_offset - self.global_offset + k, 0] = float(F1)
        self.force[g.payload_offset - self.global_offset + k, 1] = float(F2)
        st = g.cstate_range(k, 1)
        _lib.check(self._fn("csf_agent_advance")(
            _lib.MODEL_IDS[g.model], C.byref(st), C.byref(g.cparams(self.q_scale, self.q_origin)), self._off(self.force, g),
            _ptr(self.payload), self._stream()), "csf_agent_advance")
        self.gpu_launches += 1

    def source_field(self, src_xypsi, params, x, y, psi):
        """Vehicle.calcRepulsiveForce(x, y, psi): field of one source at arbitrary targets,
        without the field-of-view mask (the reference applies the mask by selecting targets)."""
        x = np.asarray(x, dtype=float)
        shape = x.shape
        xs = np.r_[float(src_xypsi[0]), x.ravel()]
        ys = np.r_[float(src_xypsi[1]), np.asarray(y, dtype=float).ravel()]
        ps = np.r_[float(src_xypsi[2]), np.broadcast_to(np.asarray(psi, dtype=float), shape).ravel()]
        n = xs.shape[0]
        dev = self.device
        (ox, oy), ext = payload_frame([np.c_[xs, ys]], margin_abs=10.0)
        q = choose_q_scale(ext)
        xd, yd, pd = (torch.as_tensor(a, dtype=torch.float64, device=dev) for a in (xs, ys, ps))
        pay = torch.zeros((n, 4), dtype=self.payload.dtype, device=dev)
        st = self._stream()
        _lib.check(self._fn("csf_pack_xypsi")(_ptr(xd), _ptr(yd), _ptr(pd), n, q, ox, oy, _ptr(pay), st), "csf_pack_xypsi")
        fp = params.to_field_params(q, False)
        fp.hfov = 2 * math.pi
        out = torch.zeros((n - 1, 2), dtype=self.dtype, device=dev)
        wsb = int(self.lib.csf_pair_workspace_bytes(1, n - 1, 4 if self.f32 else 8))
        ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
        tgt = C.c_void_p(pay.data_ptr() + self.elem_bytes)
        _lib.check(self._fn("csf_pair_forces")(_ptr(pay), 1, tgt, n - 1, C.byref(fp), _ptr(out), 0, _ptr(ws),
                                               ws.numel(), st), "csf_pair_forces")
        self.gpu_launches += 3
        o = out.to(torch.float64).cpu().numpy()
        return o[:, 0].reshape(shape), o[:, 1].reshape(shape)
