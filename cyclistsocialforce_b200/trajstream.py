"""Trajectory stream and SUMO pose batch.

The reference keeps the history of every road user in per-vehicle arrays that it writes from Python once per
vehicle and step (``traj[:, i] = s``, vehicle.py:320-325, :1407-1413; ``trajF``), and hands positions to
SUMO with one ``traci.vehicle.moveToXY`` call per vehicle and step (intersection.py:660-688).  Here:

* ``TrajectoryStream``: after every step ONE launch (``csf_copy_segments``) appends the state columns of all
  model groups and the total forces to a slot of a device ring; when a chunk of steps is full, one
  device-to-host copy on a side stream moves it into pinned memory while the simulation runs on (two
  chunks alternate).  ``drain()`` returns the recorded steps as numpy arrays.
* ``sumo_poses``: {x, y, SUMO angle} of every road user of a group after one kernel and ONE
  device-to-host copy -- the input of a batched ``moveToXY`` loop on the host.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class TrajectoryStream:
    """Records the crowd of ``engine`` step by step: call ``append()`` after every ``engine.step()``.

    chunk_steps  steps per device-to-host copy (default: as many as fit ``max_chunk_bytes``, at most 256)
    """

    def __init__(self, engine, chunk_steps=None, max_chunk_bytes=128 << 20):
        self.engine = engine
        self.device = engine.device
        self.lib = _lib.load()
        self.groups = list(engine.groups)
        if 1 + len(self.groups) > _lib.MAX_COPY_SEGMENTS:
            raise ValueError("too many model groups for one trajectory stream")
        self._layout = []                               # (group, offset in the slot, bytes)
        off = 0
        for g in self.groups:
            nb = g.state_slab.numel()
            self._layout.append((g, off, nb))
            off += (nb + 255) // 256 * 256
        self._force_off = off
        self._force_bytes = engine.force.numel() * engine.force.element_size()
        off += (self._force_bytes + 255) // 256 * 256
        self.slot_bytes = off
        if chunk_steps is None:
            chunk_steps = max(1, min(256, max_chunk_bytes // max(off, 1)))
        self.chunk_steps = int(chunk_steps)
        self._ring = [torch.empty(self.chunk_steps * off, dtype=torch.uint8, device=self.device) for _ in range(2)]
        self._host = [torch.empty(self.chunk_steps * off, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._copied = [None, None]                      # event: the D2H copy of this half has finished
        self._pending = []                               # (half, steps) in flight or not yet decoded
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._half, self._fill = 0, 0
        self._steps_total = 0
        self.launches = 0
        self._out = []                                   # decoded chunks

    # ---- device side -------------------------------------------------------------------------------
    def append(self):
        """Record the crowd's current state (the result of the step just taken) and total forces."""
        eng = self.engine
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device)
            if self._fill == 0 and self._copied[self._half] is not None:
                st.wait_event(self._copied[self._half])              # the half is being read by an older copy
            base = self._ring[self._half].data_ptr() + self._fill * self.slot_bytes
            segs = _lib.CsfCopySegments()
            for i, (g, off, nb) in enumerate(self._layout):
                segs.seg[i].src, segs.seg[i].dst, segs.seg[i].bytes = g.state_slab.data_ptr(), base + off, nb
            k = len(self._layout)
            segs.seg[k].src, segs.seg[k].dst, segs.seg[k].bytes = eng.force.data_ptr(), base + self._force_off, self._force_bytes
            segs.n = k + 1
            _lib.check(self.lib.csf_copy_segments(C.byref(segs), C.c_void_p(st.cuda_stream)), "csf_copy_segments")
            self.launches += 1
            self._fill += 1
            self._steps_total += 1
            if self._fill == self.chunk_steps:
                self._ship()

    def _ship(self):
        """Start the device-to-host copy of the open half and switch to the other one."""
        if self._fill == 0:
            return
        h, steps = self._half, self._fill
        for ph, _ in self._pending:
            if ph == h:
                self._decode_ready()                     # the host half is still undecoded: decode it first
                break
        done = torch.cuda.Event()
        filled = torch.cuda.Event()
        filled.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(filled)
            nb = steps * self.slot_bytes
            self._host[h][:nb].copy_(self._ring[h][:nb], non_blocking=True)
            done.record(self._copy_stream)
        self._copied[h] = done
        self._pending.append((h, steps))
        self._half, self._fill = 1 - h, 0

    # ---- host side ---------------------------------------------------------------------------------
    def _decode_ready(self):
        while self._pending:
            h, steps = self._pending.pop(0)
            self._copied[h].synchronize()
            buf = self._host[h][:steps * self.slot_bytes].numpy().reshape(steps, self.slot_bytes)
            layouts = [(off, {name: (o, n, np.float64 if dt == torch.float64 else np.float32)
                              for name, (o, n, dt) in g.state_layout.items()}) for g, off, nb in self._layout]
            ft = np.float64 if self.engine.force.dtype == torch.float64 else np.float32
            self._out.append(decode_chunk(buf, layouts, self._force_off, self._force_bytes, ft))

    def drain(self):
        """Everything recorded so far and not yet drained: list of chunks
        {"steps": k, "groups": [{column: (k, n) array} per model group], "force": (k, n_agents, 2)}."""
        with torch.cuda.device(self.device):
            self._ship()
            self._decode_ready()
        out, self._out = self._out, []
        return out

    @property
    def steps_recorded(self):
        return self._steps_total


def decode_chunk(buf, group_layouts, force_off, force_bytes, force_dtype):
    """One drained chunk of the ring -> arrays.  ``buf``: (steps, slot_bytes) uint8; ``group_layouts``: per model
    group (offset of its state slab in the slot, {column: (offset in the slab, n, numpy dtype)});
    returns {"steps": k, "groups": [{column: (k, n) array}], "force": (k, n_agents, 2) array}."""
    steps = buf.shape[0]
    rec = {"steps": steps, "groups": [], "force": None}
    for off, cols_layout in group_layouts:
        cols = {}
        for name, (o, n, npdt) in cols_layout.items():
            w = n * np.dtype(npdt).itemsize
            cols[name] = buf[:, off + o: off + o + w].copy().view(npdt).reshape(steps, n)
        rec["groups"].append(cols)
    rec["force"] = buf[:, force_off: force_off + force_bytes].copy().view(force_dtype).reshape(steps, -1, 2)
    return rec


def sumo_poses(group, out_dev=None, out_host=None):
    """(n, 3) array {x, y, SUMO angle in degrees} of a model group: one kernel, one device-to-host copy
    (reference intersection.py:679-688 + utils.py:119-121, per vehicle)."""
    lib = _lib.load()
    n = group.n
    with torch.cuda.device(group.device):
        if out_dev is None:
            out_dev = torch.empty((n, 3), dtype=torch.float64, device=group.device)
        fn = lib.csf_sumo_pose_f32 if group.dtype == torch.float32 else lib.csf_sumo_pose_f64
        st = torch.cuda.current_stream(group.device)
        _lib.check(fn(group.x.data_ptr(), group.y.data_ptr(), group.psi.data_ptr(), n, out_dev.data_ptr(),
                      C.c_void_p(st.cuda_stream)), "csf_sumo_pose")
        if out_host is None:
            return out_dev.cpu().numpy()
        out_host.copy_(out_dev, non_blocking=True)
        st.synchronize()
        return out_host.numpy()
