"""``SocialForceIntersection`` and the road elements with the reference's interface
(reference src/cyclistsocialforce/intersection.py), stepping on the GPU.

``step()`` issues the kernel sequence of ``engine.Engine.step`` -- nothing in this
module computes forces or dynamics on the host.  Road geometry construction
(``StraightRoadSegment`` / ``CurvedRoadSegment`` vertices) is host-side set-up and
follows the reference formulas so that the same vertices reach the road-force kernel.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .engine import AgentGroup, Engine, ObstacleGroup
from .parameters import RoadElementParameters
from .vehicle import UncontrolledVehicle, Vehicle


# ---------------------------------------------------------------------------------------
# road elements (reference intersection.py:32-250)
# ---------------------------------------------------------------------------------------
class RoadEdge:
    """Polyline exerting a repulsive force on road users (reference :214-250)."""

    def __init__(self, vertices, params=None):
        self.vertices = np.ascontiguousarray(np.asarray(vertices, dtype=float))
        self.params = params if params is not None else RoadElementParameters()

    def edges_flat(self):
        return [(self.vertices, self.params.F_0, self.params.sigma)]

    def calcRepulsiveForce(self, x, y):
        """reference :226-242, evaluated by the road-force kernel."""
        return _road_force_on_device(self.edges_flat(), x, y)


def _place(local_xy, pose_xy, angle):
    """Polyline given in a segment's own frame -> world frame: rotate by ``angle``, move to ``pose_xy``."""
    c, s_ = np.cos(angle), np.sin(angle)
    lx, ly = local_xy[:, 0], local_xy[:, 1]
    return np.column_stack((pose_xy[0] + c * lx - s_ * ly, pose_xy[1] + s_ * lx + c * ly))


class RoadSegment:
    """A piece of road with a left and a right edge (reference :72-115).  Subclasses describe their centre
    line in the segment's own frame through ``_edge(offset)`` -- the polyline ``offset`` metres to the
    left of the centre line, sampled every ``ds`` -- and ``_end_pose()``; the base class places both edges
    and the end pose ``x1`` (start pose of a following segment) in the world."""

    def __init__(self, x0, width, ds=0.1, params=None):
        self.params = RoadElementParameters()       # (the reference keeps defaults here, :73-74)
        self.x0 = x0
        self.x1 = x0
        self.width = width
        self.edges = []
        self.ds = ds

    def _build(self, frame_angle, params):
        x0 = np.asarray(self.x0, dtype=float)
        for side in (-1.0, 1.0):                     # right edge first, then left (reference order)
            self.edges.append(RoadEdge(_place(self._edge(side * self.width / 2), x0[:2], frame_angle), params=params))
        end_xy, end_heading = self._end_pose()
        self.x1 = np.array([*_place(np.array([end_xy]), x0[:2], frame_angle)[0], x0[2] + end_heading])

    def edges_flat(self):
        return [e.edges_flat()[0] for e in self.edges]

    def calcRepulsiveForce(self, x, y):
        return _road_force_on_device(self.edges_flat(), x, y)


class StraightRoadSegment(RoadSegment):
    """Straight segment of ``length`` starting at pose ``x0`` = (x, y, heading) (reference :118-146): own
    frame = x along the road."""

    def __init__(self, x0, width, length, ds=0.1, params=None):
        params = params if params is not None else RoadElementParameters()
        RoadSegment.__init__(self, x0, width, ds, params)
        self.length = length
        self._build(np.asarray(x0, dtype=float)[2], params)

    def _edge(self, offset):
        along = np.arange(0, self.length + self.ds, self.ds)
        return np.column_stack((along, np.full_like(along, offset)))

    def _end_pose(self):
        return (self.length, 0.0), 0.0


class CurvedRoadSegment(RoadSegment):
    """Circular arc of centre-line ``radius`` turning by ``angle`` to the "left" or "right" (reference
    :149-211): own frame = y along the initial heading, centre of the arc on the x axis."""

    def __init__(self, x0, width, radius, angle, direction, ds=0.1, params=None):
        params = params if params is not None else RoadElementParameters()
        RoadSegment.__init__(self, x0, width, ds, params)
        assert direction in ("left", "right"), f'direction has to be "left" or "right, instead it was {direction}'
        self.length = radius * angle
        self.radius = radius
        self.angle = angle
        self.direction = direction
        self._turn = 1.0 if direction == "left" else -1.0
        self._build(np.asarray(x0, dtype=float)[2] - np.pi / 2, params)

    def _arc(self, r, phi):
        """Points at arc angle ``phi`` on the circle of radius ``r`` about the turn centre."""
        return np.column_stack((self._turn * (r * np.cos(phi) - self.radius), r * np.sin(phi)))

    def _edge(self, offset):
        # an edge `offset` to the left of the centre line lies inside a left turn, outside a right turn
        r = self.radius - self._turn * offset
        return self._arc(r, np.linspace(0, self.angle, int(r * self.angle / self.ds)))

    def _end_pose(self):
        return tuple(self._arc(self.radius, np.array([self.angle]))[0]), self._turn * self.angle


class RoadSegmentCollection:
    """reference :32-69."""

    def __init__(self, segs):
        self.segs = segs

    def edges_flat(self):
        return [e for seg in self.segs for e in seg.edges_flat()]

    def calcRepulsiveForce(self, x, y):
        return _road_force_on_device(self.edges_flat(), x, y)

    def get_destinations_from_segments(self):
        return [seg.x1[0] for seg in self.segs], [seg.x1[1] for seg in self.segs]

    def __getitem__(self, i):
        if not isinstance(i, int):
            raise ValueError("Subscription index must be integer!")
        return self.segs[i]


def _road_force_on_device(edges, x, y):
    """Evaluate the road force of ``edges`` at arbitrary points through the C ABI (f64)."""
    lib = _lib.load()
    x = np.asarray(x, dtype=float)
    shape = x.shape
    xd = torch.as_tensor(np.ascontiguousarray(x.ravel()), device="cuda")
    yd = torch.as_tensor(np.ascontiguousarray(np.asarray(y, dtype=float).ravel()), device="cuda")
    out = torch.zeros((xd.numel(), 2), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for k, (verts, F_0, sigma) in enumerate(edges):
        vd = torch.as_tensor(np.ascontiguousarray(verts), device="cuda").contiguous()
        _lib.check(lib.csf_road_forces_f64(xd.data_ptr(), yd.data_ptr(), xd.numel(), vd.data_ptr(), vd.shape[0],
                                           float(F_0), float(sigma), out.data_ptr(), 1 if k else 0, st),
                   "csf_road_forces_f64")
    o = out.cpu().numpy()
    return o[:, 0].reshape(shape), o[:, 1].reshape(shape)


# ---------------------------------------------------------------------------------------
# the interaction manager
# ---------------------------------------------------------------------------------------
class SocialForceIntersection:
    """Manages a social force intersection or open space (reference :253-915).

    Extra keyword arguments of this implementation (all optional):
      dtype        torch.float32 (production build, default) or torch.float64 (verification)
      record_traj  keep the per-vehicle ``traj`` / ``trajF`` / ``F`` histories on the host; default: on for
                   <= 64 road users (one device->host copy per step).  ``record_traj=True`` on a larger
                   crowd records through a device ring that the host drains a chunk of steps at a time
                   (trajstream.TrajectoryStream: one launch per step, one copy per chunk); the per-vehicle
                   arrays are filled in when ``vehicle.traj`` is read or ``flush_trajectories()`` is called.
      sumo_client  a traci-like object (``.vehicle.moveToXY``); with ``activate_sumo_cosimulation=True``
                   every step ends with a batched position update (one kernel + one device->host copy,
                   then the client calls; reference :660-688).  Without a client ``traci`` is imported.
    """

    def __init__(self, vehicleList, id="", priority_rule="unregulated", animate=False, axes=None,
                 activate_sumo_cosimulation=False, net=None, road_elements=(), bicycle_drawing_kwargs=None,
                 dtype=torch.float32, record_traj=None, device="cuda", sumo_client=None, traj_chunk_steps=None,
                 _private=False):
        if animate:
            raise NotImplementedError("matplotlib animation is outside the accelerated stepping path")
        if activate_sumo_cosimulation and sumo_client is None:
            try:
                import traci as sumo_client          # noqa: F811  (the reference's client, if installed)
            except ImportError as e:
                raise ImportError("activate_sumo_cosimulation=True needs the SUMO python client (traci) or a "
                                  "sumo_client object with .vehicle.moveToXY") from e
        self._sumo = sumo_client if activate_sumo_cosimulation else None
        assert isinstance(id, str), "Intersection ID has to be a string."
        assert priority_rule in ("p2r", "unregulated"), "Priority rule has to be one of ('p2r','unregulated')"
        self.id = id
        self.priority_rule = priority_rule
        self.animate = False
        self.ax = axes
        self.activate_sumo_cosimulation = bool(activate_sumo_cosimulation)
        self._traj_stream = None
        self._traj_chunk_steps = traj_chunk_steps
        self.road_elements = list(road_elements)
        self.is_first_step = True
        self.hist_n_vecs = []
        self.dtype = dtype
        self.device = device
        self._record_traj_opt = record_traj
        self._private = _private
        self.vehicles = list(vehicleList)
        self._engine = None
        self._rebuild()

    # ---- construction of the device crowd ---------------------------------------------------------
    def _rebuild(self):
        """(Re)create the device crowd from the vehicle list; per-vehicle device records are
        carried over (add_road_user / remove_road_user, reference :458-539, :576-634)."""
        if self._engine is not None:
            self._pull_records()
        self.n_bikes = len(self.vehicles)
        controlled = [v for v in self.vehicles if v.MODEL is not None]
        obstacles = [v for v in self.vehicles if v.MODEL is None]
        # one group per (model, params object identity-equivalent field/agent parameters)
        groups, order = [], {}
        for v in controlled:
            key = (v.MODEL, _params_key(v.params))
            order.setdefault(key, []).append(v)
        for (model, _), vs in order.items():
            s0 = np.stack([v._s for v in vs])
            g = AgentGroup(model, s0, vs[0].params, vd_default=[v.params.v_desired_default for v in vs],
                           destqueues=[v._destqueue for v in vs], dtype=self.dtype, device=self.device,
                           stream_ids=[v._stream_id for v in vs])
            for k, v in enumerate(vs):
                if v._record is not None:
                    g.import_record(k, v._record)
                v._owner, v._group, v._k = self, g, k
            groups.append(g)
        obs = []
        if obstacles:
            pk = {}
            for v in obstacles:
                pk.setdefault(_params_key(v.params), []).append(v)
            for vs in pk.values():
                o = ObstacleGroup(np.stack([v._s[:3] for v in vs]), vs[0].params, device=self.device)
                o.vehicles = vs
                for v in vs:
                    v._owner, v._group, v._k = self, None, -1
                obs.append(o)
        self._groups, self._obstacle_groups = groups, obs
        self._make_engine()

    def _make_engine(self):
        """(Re)bind the engine to the current device groups."""
        self.n_bikes = len(self.vehicles)
        self._obstacles_dirty = False
        edges = [e for el in self.road_elements for e in el.edges_flat()]
        if self.n_bikes > 0:
            # (the engine being replaced hands its device buffers over: churn does not go through the allocator)
            self._engine = Engine(self._groups, obstacles=self._obstacle_groups, priority_rule=self.priority_rule,
                                  road_edges=edges, dtype=self.dtype, device=self.device, reuse=self._engine)
        else:
            self._engine = None
        self._invalidate()
        rt = self._record_traj_opt
        self.record_traj = (self.n_bikes <= 64) if rt is None else bool(rt)
        self.flush_trajectories()                 # (records of the previous binding, if any)
        self._traj_stream = None
        if self.record_traj and self.n_bikes > 64 and self._engine is not None:
            from .trajstream import TrajectoryStream
            self._traj_stream = TrajectoryStream(self._engine, chunk_steps=self._traj_chunk_steps)
            self._traj_groups = list(self._groups)

    # ---- churn without a host round trip of the crowd (reference :458-539, :576-634) -------------------
    def _device_add(self, user):
        """Append one controlled road user to the device group of its (model, parameter set): the
        group's state stays on the device (AgentGroup.concat).  False if there is no such group."""
        if self._engine is None or user.MODEL is None:
            return False
        key = (user.MODEL, _params_key(user.params))
        for gi, g in enumerate(self._groups):
            if (g.model, _params_key(g.params)) == key:
                break
        else:
            return False
        one = AgentGroup(user.MODEL, np.asarray(user._s, float)[None, :], g.params,
                         vd_default=[user.params.v_desired_default], destqueues=[user._destqueue],
                         dtype=self.dtype, device=self.device, stream_ids=[user._stream_id])
        if user._record is not None:
            one.import_record(0, user._record)
        new = AgentGroup.concat(g, one)
        for v in self.vehicles:
            if v._group is g:
                v._group = new
        user._owner, user._group, user._k = self, new, g.n
        self._groups[gi] = new
        self._make_engine()
        return True

    def _device_remove(self, removed):
        """Drop controlled road users from their device groups (AgentGroup.select); only the removed
        users' records travel to the host.  False if an obstacle is among them (full rebuild)."""
        if self._engine is None or any(v._group is None for v in removed):
            return False
        by_group = {}
        for v in removed:
            by_group.setdefault(id(v._group), []).append(v)
        groups = []
        for g in self._groups:
            vs = by_group.get(id(g))
            if not vs:
                groups.append(g)
                continue
            rm = sorted(v._k for v in vs)
            recs = g.select(rm).export_records()
            for v in vs:
                r = recs[rm.index(v._k)]
                v._record, v._s = r, r["_s"]
                v._i, v._destpointer = int(r["step_i"]), int(r["dest_ptr"])
            gone = set(rm)
            keep = [k for k in range(g.n) if k not in gone]
            newk = {k: j for j, k in enumerate(keep)}
            new = g.select(keep) if keep else None
            for v in self.vehicles:
                if v._group is g and v._k in newk:
                    v._group, v._k = new, newk[v._k]
            if new is not None:
                groups.append(new)
        for v in removed:
            v._owner = v._group = None
            v._k = -1
        self._groups = groups
        self._make_engine()
        return True

    def _pull_records(self):
        for g in self._groups:
            recs = g.export_records()
            for v in self.vehicles:
                if v._group is g:
                    v._record = recs[v._k]
                    v._s = recs[v._k]["_s"]
                    v._i = int(recs[v._k]["step_i"])
                    v._destpointer = int(recs[v._k]["dest_ptr"])

    # ---- cached host snapshots -----------------------------------------------------------------------
    def _invalidate(self):
        self._cache = {}

    def _host_state(self, g):
        key = ("s", id(g))
        if key not in self._cache:
            self._cache[key] = g.states_numpy()
        return self._cache[key]

    def _host_field(self, g, name):
        key = (name, id(g))
        if key not in self._cache:
            self._cache[key] = getattr(g, name).cpu().numpy()
        return self._cache[key]

    def _host_force(self):
        if "force" not in self._cache:
            self._cache["force"] = self._engine.force.to(torch.float64).cpu().numpy()
        return self._cache["force"]

    # ---- reference attributes ------------------------------------------------------------------------------
    def _xypsi(self):
        out = np.zeros((self.n_bikes, 3))
        for i, v in enumerate(self.vehicles):
            out[i] = v.s[:3]
        return out

    @property
    def vehicleX(self):
        return self._xypsi()[:, 0:1]

    @property
    def vehicleY(self):
        return self._xypsi()[:, 1:2]

    @property
    def vehicleTheta(self):
        return self._xypsi()[:, 2:3]

    def update_road_user_positions(self):
        """reference :660-688: state -> pair payload (device side); with SUMO co-simulation the positions go
        to the client -- one kernel and one device->host copy per model group, then one ``moveToXY`` per
        road user from the batch."""
        if self._engine is not None:
            self._sync_obstacles()
            self._engine.pack()
            if self._sumo is not None:
                self._push_to_sumo()

    def _push_to_sumo(self):
        from .trajstream import sumo_poses
        for g in self._groups:
            poses = sumo_poses(g)
            for v in self.vehicles:
                if v._group is g:
                    x, y, ang = poses[v._k]
                    self._sumo.vehicle.moveToXY(v.id, "", -1, float(x), float(y), angle=float(ang), keepRoute=6)

    def flush_trajectories(self):
        """Move everything the trajectory stream has recorded into the vehicles' ``traj`` / ``trajF`` /
        ``F`` histories (crowds of more than 64 road users with ``record_traj=True``)."""
        ts = getattr(self, "_traj_stream", None)
        if ts is None:
            return
        chunks = ts.drain()
        if not chunks:
            return
        by_group = {id(g): [v for v in self.vehicles if v._group is g] for g in self._traj_groups}
        step_now = {id(g): g.step_i.cpu().numpy() for g in self._traj_groups}
        total = sum(c["steps"] for c in chunks)
        done = 0
        for c in chunks:
            k = c["steps"]
            for gi, g in enumerate(self._traj_groups):
                cols = c["groups"][gi]
                names = [n for n in ("x", "y", "psi", "v", "delta", "theta", "deltadot", "thetadot") if n in cols]
                S = np.stack([cols[n].astype(float) for n in names], axis=1)            # (k, n_states, n)
                off = g.payload_offset - self._engine.global_offset
                for v in by_group[id(g)]:
                    L = v._traj.shape[1]
                    last = int(step_now[id(g)][v._k]) - (total - done - k)                # step index of row k-1
                    idx = (np.arange(last - k + 1, last + 1)) % L
                    v._traj[:, idx] = S[:, :v._traj.shape[0], v._k].T
                    f = c["force"][:, off + v._k, :].astype(float)
                    v.F.extend(np.hypot(f[:, 0], f[:, 1]).tolist())
                    if v.trajF is not None:
                        v.trajF[:, idx] = f.T
            done += k

    def _sync_obstacles(self):
        if self._obstacles_dirty:
            for o in self._obstacle_groups:
                o.set(np.stack([v._s[:3] for v in o.vehicles]))
            self._engine.pack()
            self._obstacles_dirty = False

    # ---- road users --------------------------------------------------------------------------------------
    def add_road_user(self, user):
        """reference :458-539."""
        self.flush_trajectories()                 # (before the device groups are re-bound)
        self.vehicles.append(user)
        if not self._device_add(user):
            self._rebuild()

    def get_road_user_ids(self):
        return [v.id for v in self.vehicles]

    def remove_road_user(self, i):
        """reference :576-616."""
        self.flush_trajectories()
        v = self.vehicles.pop(i)
        if not self._device_remove([v]):
            self.vehicles.insert(i, v)
            self._pull_records()
            self.vehicles.pop(i)
            v._owner = v._group = None
            self._rebuild()
        return v

    def remove_road_users_by_id(self, ids):
        """reference :618-634."""
        self.flush_trajectories()
        gone = [v for v in self.vehicles if v.id in ids]
        keep = [v for v in self.vehicles if v.id not in ids]
        self.vehicles = keep
        if not self._device_remove(gone):
            self.vehicles = keep + gone
            self._pull_records()
            self.vehicles = keep
            for v in gone:
                v._owner = v._group = None
            self._rebuild()

    # ---- the hot path ---------------------------------------------------------------------------------------
    def _force_in_vehicle_order(self):
        f = self._host_force()
        out = np.zeros((self.n_bikes, 2))
        for i, v in enumerate(self.vehicles):
            if v._group is not None:
                out[i] = f[v._group.payload_offset + v._k]
        return out

    def calc_forces(self):
        """reference :747-864 -> (Fx, Fy) ndarrays of shape (n_bikes,)."""
        if self._engine is None:
            return np.zeros(0), np.zeros(0)
        self._sync_obstacles()
        self._engine.calc_forces()
        self._invalidate()
        f = self._force_in_vehicle_order()
        if self.record_traj:
            for i, v in enumerate(self.vehicles):
                v.F.append(float(np.hypot(f[i, 0], f[i, 1])))
        return f[:, 0].copy(), f[:, 1].copy()

    def step(self):
        """reference :866-896."""
        self.is_first_step = False
        if self.n_bikes > 0 and self._engine is not None:
            self._sync_obstacles()
            self._engine.step()
            for o in self._obstacle_groups:
                for v in o.vehicles:
                    v.step()
            self._invalidate()
            if self._traj_stream is not None:
                self._traj_stream.append()
            elif self.record_traj:
                self._record_histories()
            if self._sumo is not None:
                self._push_to_sumo()
        self.hist_n_vecs.append(self.n_bikes)

    def _record_histories(self):
        f = self._force_in_vehicle_order()
        for i, v in enumerate(self.vehicles):
            if v._group is None:
                continue
            s = self._host_state(v._group)[v._k]
            idx = int(self._host_field(v._group, "step_i")[v._k]) % v._traj.shape[1]
            v._traj[:, idx] = s
            v.F.append(float(np.hypot(f[i, 0], f[i, 1])))
            if v.trajF is not None:
                v.trajF[:, idx] = f[i]
        self._engine.check_status()

    def check_status(self):
        """Raise if the device flagged non-finite values, an invalid navigation state or a
        position outside the payload range (the reference raises from Python in these cases)."""
        if self._engine is not None:
            return self._engine.check_status()
        return []

    def set_animated(self, animated):
        pass

    # ---- per-vehicle operations (Vehicle.calcDestinationForce / calcRepulsiveForce / step) ---------------------
    def _vehicle_dest_force(self, v):
        f = self._engine.agent_dest_force(v._group, v._k)
        self._invalidate()
        return f

    def _vehicle_rep_force(self, v, x, y, psi):
        return self._engine.source_field(v.s[:3], v.params, x, y, psi)

    def _vehicle_step(self, v, F1, F2):
        self._engine.agent_advance(v._group, v._k, F1, F2)
        self._invalidate()
        if self.record_traj:
            s = self._host_state(v._group)[v._k]
            idx = int(self._host_field(v._group, "step_i")[v._k]) % v._traj.shape[1]
            v._traj[:, idx] = s
            if v.trajF is not None:
                v.trajF[:, idx] = (F1, F2)


def _params_key(p):
    """Vehicles share a device group when every kernel-visible parameter agrees
    (``v_desired_default`` is per agent and therefore excluded)."""
    items = []
    for k, val in sorted(vars(p).items()):
        if k in ("_v_desired_default", "verbose", "calib_mode", "rep_force", "dest_force", "dynamics"):
            continue
        if isinstance(val, (list, tuple, np.ndarray)):
            val = tuple(np.asarray(val).ravel().tolist())
        elif isinstance(val, dict):
            val = tuple(sorted((kk, float(vv)) for kk, vv in val.items() if isinstance(vv, (int, float))))
        elif not isinstance(val, (int, float, str, bool, type(None))):
            continue
        items.append((k, val))
    return (type(p).__name__, tuple(items))
