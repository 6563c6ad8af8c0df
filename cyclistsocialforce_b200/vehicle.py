"""Vehicle classes with the reference's names and signatures (reference
src/cyclistsocialforce/vehicle.py), backed by device-resident state.

A vehicle is a thin host object: construction parameters, the destination queue and --
once it belongs to a ``SocialForceIntersection`` -- an index into that intersection's
device crowd.  ``.s``, ``.destpointer``, ``.znav``, ``.force`` read the device state
(one batched device->host copy per step, shared by all vehicles of the intersection).
A vehicle that is used on its own (``calcDestinationForce()``, ``step(Fx, Fy)``) gets a
private one-vehicle intersection on first use, so the same kernels run in every case.

Fixed reference defects: the ``TwoDBicycle``/``InvPendulumBicycle``
constructors accept their documented signature (reference vehicle.py:1359 raises
TypeError); stepping never needs a matplotlib Axes.
"""
from __future__ import annotations

import numpy as np

from .parameters import (
    BalancingRiderBicycleParameters,
    BicycleParameters,
    CarParameters,
    InvPendulumBicycleParameters,
    PlanarPointBicycleParameters,
    VehicleParameters,
)
from .utils import limitAngle


class Vehicle:
    """Parent class for all vehicle types (reference vehicle.py:49-917)."""

    MODEL = None                      # engine model name; None = not steppable on the device
    PARAMS_TYPE = VehicleParameters
    REQUIRED_PARAMS = ("d_arrived_inter", "hfov")
    N_STATES = 4
    STATE_NAMES = ["x[m]", "y[m]", "psi[rad]", "v[m/s]"]
    _n_created = 0

    def __init__(self, s0, id="unknown", route=(), saveForces=False, params=None, dest_force_func=None,
                 rep_force_func=None, uncontrolled=False, uncontrolled_traj=()):
        if params is None:
            self.params = self.PARAMS_TYPE()
        else:
            if not isinstance(params, self.PARAMS_TYPE):
                raise TypeError(f"Params must be a '{self.PARAMS_TYPE.__name__}' object. "
                                f"Instead it was '{type(params).__name__}'.")
            self.params = params
        if dest_force_func is not None or rep_force_func is not None:
            raise NotImplementedError(
                "custom Python force callbacks (rep_force_func / dest_force_func) cannot run inside the "
                "CUDA kernels; they are outside the accelerated path")
        if len(s0) < self.N_STATES:
            raise ValueError(f"The initial state s0 has to be size {self.N_STATES} with states "
                             f"{self.STATE_NAMES}. Instead it was {s0}.")
        s0 = list(s0)[:self.N_STATES]
        self._s = np.array(s0, dtype=float)
        self._s[2] = limitAngle(float(s0[2]))
        self.s_names = self.STATE_NAMES
        self._i = 0
        assert isinstance(id, str), "User ID has to be a string."
        self.id = id
        assert isinstance(route, tuple), "Route has to be a tuple"
        assert all(isinstance(r, str) for r in route), "Edge IDs in route list have to be str"
        self.follow_route = bool(route)
        self.route = route
        self.saveForces = saveForces
        self.drawing = None
        # destination queue: first entry is the start position (reference :183-185)
        self._destqueue = np.array([[self._s[0], self._s[1], 0.0]])
        self._destpointer = 0
        self._znav = np.array([True, False, False])
        self.F = []
        self.uncontrolled = bool(uncontrolled)
        n_traj = int(30 / self.params.t_s)
        self._traj = np.zeros((self.N_STATES, n_traj))
        self._traj[:, 0] = self._s
        self.trajF = np.zeros((2, n_traj)) if saveForces else None
        # binding to a device crowd
        self._owner = None     # SocialForceIntersection
        self._group = None     # engine.AgentGroup
        self._k = -1
        self._record = None    # full per-agent device record kept across re-binding
        # key of this road user's random stream (stochastic rider behaviour): its serial number in the process
        self._stream_id = Vehicle._n_created
        Vehicle._n_created += 1

    # ---- state views -------------------------------------------------------------------------
    def _bound(self):
        return self._owner is not None and self._group is not None

    @property
    def s(self):
        if self._bound():
            return self._owner._host_state(self._group)[self._k].copy()
        return self._s

    @s.setter
    def s(self, value):
        value = np.asarray(value, dtype=float)
        self._s = value.copy()
        if self._bound():
            self._group.set_state_row(self._k, value)
            self._owner._invalidate()

    @property
    def i(self):
        if self._bound():
            return int(self._owner._host_field(self._group, "step_i")[self._k])
        return self._i

    @property
    def destpointer(self):
        if self._bound():
            return int(self._owner._host_field(self._group, "dest_ptr")[self._k])
        return self._destpointer

    @property
    def destqueue(self):
        return self._destqueue

    @property
    def dest(self):
        return self._destqueue[min(self.destpointer, len(self._destqueue) - 1)]

    @property
    def znav(self):
        if self._bound():
            z = int(self._owner._host_field(self._group, "znav")[self._k])
            return np.array([bool(z & 1), bool(z & 2), bool(z & 4)])
        return self._znav

    @property
    def force(self):
        if self._bound():
            f = self._owner._host_force()[self._group.payload_offset + self._k]
            return (float(f[0]), float(f[1]))
        return (0.0, 0.0)

    @property
    def traj(self):
        if self._owner is not None:
            self._owner.flush_trajectories()          # (no-op unless a trajectory stream is recording)
        return self._traj

    # ---- destinations --------------------------------------------------------------------------
    def isLastDest(self):
        """reference :537-543."""
        return self.destpointer + 1 >= self._destqueue.shape[0]

    def getDestinationDistance(self):
        """reference :596-604."""
        d, s = self.dest, self.s
        return float(np.sqrt((d[0] - s[0]) ** 2 + (d[1] - s[1]) ** 2))

    def setDestinations(self, x, y, stop=None, reset=False):
        """reference :606-647."""
        x = np.array([x], dtype=float).flatten()
        y = np.array([y], dtype=float).flatten()
        stop = np.zeros_like(x) if stop is None else np.array([stop], dtype=float).flatten()
        if reset or self._destqueue is None:
            self._destqueue = np.c_[x, y, stop]
            self._set_destpointer(0)
        else:
            self._destqueue = np.vstack((self._destqueue, np.c_[x, y, stop]))
        self._push_destqueue()

    def setSplineDestinations(self, x, y, npoints, stop=False, reset=False):
        """reference :649-693 (host-side set-up; uses scipy FITPACK like the reference)."""
        from scipy import interpolate
        assert len(x) >= 3, "Provide at least 3 points to calculate a cubic trajectory prototype"
        s = self.s
        x = np.insert(np.array(x, dtype=float), 0, s[0])
        y = np.insert(np.array(y, dtype=float), 0, s[1])
        tck, _ = interpolate.splprep((x, y), s=0.0)
        x_i, y_i = interpolate.splev(np.linspace(0, 1, npoints), tck)
        if stop:
            st = np.zeros_like(x_i)
            st[-1] = 1.0
            self.setDestinations(x_i, y_i, stop=st, reset=reset)
        else:
            self.setDestinations(x_i, y_i, reset=reset)

    def stop(self, stoptype=0, stopdest=None):
        """reference :459-503.  Only stoptype 0 (stop at the next destination) is supported:
        stoptype 1 reads a non-existent ``params.AMAX`` in the reference (:486) and
        stoptype 2 is overwritten by the next ``updateDestination`` (:586)."""
        if stoptype == 0:
            # reference: self.dest is a view into destqueue, so this sets the queue's stop flag
            self._destqueue[min(self.destpointer, len(self._destqueue) - 1), 2] = 1.0
            self._push_destqueue()
        elif stoptype in (1, 2):
            raise NotImplementedError("stop types 1 and 2 are broken in the reference and not supported")
        else:
            raise ValueError("Stop type has to be one of [0,1,2].")

    def go(self, gotype=0):
        """reference :505-535."""
        if gotype == 0:
            self._destqueue[min(self.destpointer, len(self._destqueue) - 1), 2] = 0.0
            self._push_destqueue()
        elif gotype != 1:
            raise ValueError("Go type has to be one of [0,1].")
        self.arrived = False

    def _set_destpointer(self, p):
        self._destpointer = p
        if self._bound():
            self._group.dest_ptr[self._k] = p
            self._owner._invalidate()

    def _push_destqueue(self):
        if self._bound():
            self._group.set_destqueue(self._k, self._destqueue)
            self._owner._invalidate()

    # ---- per-vehicle operations (reference plugin hooks) -------------------------------------------
    def _ensure_owner(self):
        if self._owner is None:
            from .intersection import SocialForceIntersection
            SocialForceIntersection([self], id="__private__", _private=True)
        return self._owner

    def calcDestinationForce(self):
        """reference :281-299 / per-class overrides.  Mutates the destination pointer and the
        navigation state like the reference does."""
        if self.MODEL is None:
            return 0, 0
        return self._ensure_owner()._vehicle_dest_force(self)

    def calcRepulsiveForce(self, x, y, psi):
        """Force this vehicle exerts on road users at (x, y) with headings psi
        (reference :250-279, TwoDBicycle.calcRepulsiveForce :1560-1648)."""
        return self._ensure_owner()._vehicle_rep_force(self, x, y, psi)

    def step(self, F1=0, F2=0):
        """Advance this vehicle by one step under the force (F1, F2) (reference :301-328)."""
        if self.MODEL is None:
            self._i += 1
            return
        self._ensure_owner()._vehicle_step(self, F1, F2)

    # ---- drawing (out of scope: presentation) ----------------------------------------------------
    def add_drawing(self, ax, drawing=None, **kwargs):
        self.drawing = drawing

    def update_drawing(self, Fdest=None, Frep=None, Fres=None):
        pass


class UncontrolledVehicle(Vehicle):
    """Obstacle / externally controlled vehicle following a prescribed trajectory
    (reference :920-987).  Exerts the TwoDBicycle field on others, feels nothing."""

    MODEL = None
    PARAMS_TYPE = CarParameters

    def __init__(self, s0, trajectory=(), **kwargs):
        kwargs.setdefault("params", self.PARAMS_TYPE())
        Vehicle.__init__(self, s0, **kwargs)
        self.uncontrolled = True
        if len(trajectory) > 0:
            self._traj = np.array(trajectory, dtype=float)

    def step(self, Fx=None, Fy=None):
        self._i += 1
        if np.shape(self._traj)[1] > self._i:
            self._s = self._traj[:, self._i].copy()
        if self._owner is not None:
            self._owner._obstacles_dirty = True

    def calcDestinationForce(self):
        return 0, 0


class Bicycle(Vehicle):
    """v0.1 kinematic bicycle (reference :990-1289)."""

    MODEL = "bicycle"
    PARAMS_TYPE = BicycleParameters
    N_STATES = 5
    STATE_NAMES = ["x[m]", "y[m]", "psi[rad]", "v[m/s]", "delta[rad]"]

    def __init__(self, s0, **kwargs):
        Vehicle.__init__(self, s0, **kwargs)


class TwoDBicycle(Bicycle):
    """"2D model" of Schmidt et al. 2023 (reference :1292-1648)."""

    MODEL = "twod"
    PARAMS_TYPE = InvPendulumBicycleParameters

    def __init__(self, s0, id="unknown", route=(), saveForces=False, params=None):
        Vehicle.__init__(self, s0, id=id, route=route, saveForces=saveForces, params=params)


class InvPendulumBicycle(TwoDBicycle):
    """"Inverted pendulum model" of Schmidt et al. 2023 (reference :1651-1950)."""

    MODEL = "invpendulum"
    PARAMS_TYPE = InvPendulumBicycleParameters
    N_STATES = 6
    STATE_NAMES = ["x[m]", "y[m]", "psi[rad]", "v[m/s]", "delta[rad]", "theta[rad]"]

    def __init__(self, s0, **kwargs):
        Vehicle.__init__(self, s0, **kwargs)

    @property
    def zrid(self):
        if self._bound():
            z = int(self._owner._host_field(self._group, "ip_zrid")[self._k])
            return np.array([bool(z & 1), bool(z & 2)])
        walk = self._s[3] < self.params.v_max_walk
        return np.array([not walk, walk])


#: BASELINE.json / north-star spelling
InvertedPendulumBicycle = InvPendulumBicycle


class BalancingRiderBicycle(Vehicle):
    """Whipple-Carvallo bicycle with full-state feedback (reference :1953-1988,
    dynamics.py:261-705)."""

    MODEL = "balancingrider"
    PARAMS_TYPE = BalancingRiderBicycleParameters
    N_STATES = 8
    STATE_NAMES = ["x[m]", "y[m]", "psi[rad]", "v[m/s]", "delta[rad]", "phi[rad]", "deltadot[rad/s]",
                   "phidot[rad/s]"]

    def __init__(self, s0, **kwargs):
        Vehicle.__init__(self, s0, **kwargs)


class PlanarPointBicycle(Vehicle):
    """Mass-less particle with first-order yaw tracking (reference :1991-2028,
    dynamics.py:802-1079)."""

    MODEL = "planarpoint"
    PARAMS_TYPE = PlanarPointBicycleParameters
    N_STATES = 4

    def __init__(self, s0, **kwargs):
        Vehicle.__init__(self, s0, **kwargs)
