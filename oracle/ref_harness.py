"""Headless import harness for the *reference's own* code (test infrastructure).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Works only where
``/root/reference`` exists (the build container); it cannot travel to the GPU
box.  Its job is (1) to pin the restated oracle (``oracle/csf_oracle.py``)
against the reference and (2) to generate the golden vectors committed under
``tests/golden/`` (``tests/golden/make_golden.py``).

What it does:

* ``sys.modules`` stubs for the presentation / SUMO packages the reference
  imports at module level but that the step path never calls: matplotlib,
  mpl_toolkits, pypaperutils, mypyutils, traci, sumolib, libsumo.
* Third-party *arithmetic* that is absent from this image is substituted by
  restatements of the published algorithms, each marked below:
    - ``control`` (python-control, unpinned in reference pyproject.toml:21):
      ``ss``/``StateSpace``, ``ctrb``, ``place`` (= scipy.signal.place_poles
      'YT', as python-control does), ``forced_response`` (first-order-hold
      matrix-exponential scheme of python-control's timeresp.py).
    - ``bicycleparameters`` (unpinned, pyproject.toml:25):
      ``Meijaard2007ParameterSet`` / ``Meijaard2007Model`` with
      ``form_reduced_canonical_matrices`` / ``form_state_space_matrices`` from
      Meijaard et al. 2007, Appendix A.
  Everything routed through these two substitutes is "parity unpinned" with
  respect to the real packages (DESIGN.md says so).
* A corrected ``TwoDBicycle.__init__`` (reference src/cyclistsocialforce/
  vehicle.py:1359 passes positionals to the keyword-only
  ``Bicycle.__init__`` at :1020 -> TypeError).  Only the constructor is
  replaced; all maths is the reference's.
"""
from __future__ import annotations

import os
import sys
import types
from unittest import mock

import numpy as np

REFERENCE_SRC = os.environ.get("CSF_REFERENCE_SRC", "/root/reference/src")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "cyclistsocialforce"))


# --------------------------------------------------------------------------
# substitute: python-control
# --------------------------------------------------------------------------
def _build_control_module() -> types.ModuleType:
    import scipy.linalg
    import scipy.signal

    ct = types.ModuleType("control")

    def _col(B):
        B = np.array(B, dtype=float)
        return B.reshape(-1, 1) if B.ndim == 1 else B

    def _row(C):
        C = np.array(C, dtype=float)
        return C.reshape(1, -1) if C.ndim == 1 else C

    class StateSpace:
        def __init__(self, A, B, C, D, *a, **k):
            self.A = np.atleast_2d(np.array(A, dtype=float))
            self.B = _col(B)
            self.C = _row(C)
            if np.isscalar(D) or np.ndim(D) == 0:
                self.D = np.full((self.C.shape[0], self.B.shape[1]), float(D))
            else:
                self.D = np.atleast_2d(np.array(D, dtype=float))

        def poles(self):
            return np.linalg.eigvals(self.A)

    class TimeResponseData:
        def __init__(self, t, y, x, return_x):
            self.time, self.outputs, self.states = t, y, x
            self._return_x = return_x

        def __iter__(self):
            if self._return_x:
                return iter((self.time, self.outputs, self.states))
            return iter((self.time, self.outputs))

    def forced_response(sys_, T=None, U=0.0, X0=0.0, transpose=False,
                        return_x=False, squeeze=None, **kw):
        A, B, C, D = sys_.A, _col(sys_.B), sys_.C, sys_.D
        n, m = A.shape[0], B.shape[1]
        T = np.asarray(T, dtype=float)
        nt = T.shape[0]
        U = np.asarray(U, dtype=float)
        if U.ndim == 0:
            U = np.full((m, nt), float(U))
        elif U.ndim == 1:
            U = U.reshape(1, -1)
        x0 = np.zeros(n)
        X0 = np.asarray(X0, dtype=float).reshape(-1)
        if X0.size == n:
            x0 = X0
        elif X0.size == 1:
            x0 = np.full(n, X0[0])
        dt = 1.0 if nt == 1 else T[1] - T[0]
        M = np.block([
            [A * dt, B * dt, np.zeros((n, m))],
            [np.zeros((m, n + m)), np.identity(m)],
            [np.zeros((m, n + 2 * m))],
        ])
        expM = scipy.linalg.expm(M)
        Ad = expM[:n, :n]
        Bd1 = expM[:n, n + m:]
        Bd0 = expM[:n, n:n + m] - Bd1
        xout = np.zeros((n, nt))
        xout[:, 0] = x0
        for i in range(1, nt):
            xout[:, i] = Ad @ xout[:, i - 1] + Bd0 @ U[:, i - 1] + Bd1 @ U[:, i]
        yout = C @ xout + D @ U
        return TimeResponseData(T, yout, xout, return_x)

    def place(A, B, p):
        res = scipy.signal.place_poles(np.array(A, float), _col(B),
                                       np.array(p), method="YT")
        return res.gain_matrix

    def ctrb(A, B):
        A = np.array(A, float)
        B = _col(B)
        n = A.shape[0]
        return np.hstack([np.linalg.matrix_power(A, i) @ B for i in range(n)])

    ct.StateSpace = StateSpace
    ct.ss = lambda *a, **k: StateSpace(*a, **k)
    ct.forced_response = forced_response
    ct.place = place
    ct.ctrb = ctrb
    ct.TimeResponseData = TimeResponseData
    return ct


# --------------------------------------------------------------------------
# substitute: bicycleparameters (Meijaard et al. 2007, Appendix A)
# --------------------------------------------------------------------------
from oracle.csf_oracle import meijaard_canonical  # noqa: E402  (restated Meijaard 2007 App. A)


def _build_bicycleparameters_modules():
    bp = types.ModuleType("bicycleparameters")
    pd_ = types.ModuleType("bicycleparameters.parameter_dicts")
    ps_ = types.ModuleType("bicycleparameters.parameter_sets")
    md_ = types.ModuleType("bicycleparameters.models")
    pd_.meijaard2007_browser_jason = {}

    class Meijaard2007ParameterSet:
        def __init__(self, parameters, includes_rider=True):
            self.parameters = dict(parameters)
            self.includes_rider = includes_rider

    class Meijaard2007Model:
        def __init__(self, parameter_set):
            self.parameter_set = parameter_set

        def form_reduced_canonical_matrices(self, **over):
            p = dict(self.parameter_set.parameters, **over)
            return meijaard_canonical(p)

        def form_state_space_matrices(self, **over):
            p = dict(self.parameter_set.parameters, **over)
            M, C1, K0, K2 = meijaard_canonical(p)
            v, g = p["v"], p["g"]
            Minv = np.linalg.inv(M)
            A = np.zeros((4, 4))
            A[0:2, 2:4] = np.eye(2)
            A[2:4, 0:2] = -Minv @ (g * K0 + v**2 * K2)
            A[2:4, 2:4] = -Minv @ (v * C1)
            B = np.zeros((4, 2))
            B[2:4, :] = Minv
            return A, B

    ps_.Meijaard2007ParameterSet = Meijaard2007ParameterSet
    md_.Meijaard2007Model = Meijaard2007Model
    bp.parameter_dicts, bp.parameter_sets, bp.models = pd_, ps_, md_
    return {"bicycleparameters": bp, "bicycleparameters.parameter_dicts": pd_,
            "bicycleparameters.parameter_sets": ps_, "bicycleparameters.models": md_}


# --------------------------------------------------------------------------
# stubs for presentation / SUMO packages
# --------------------------------------------------------------------------
_STUBS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.axes", "matplotlib.patches",
    "matplotlib.collections", "matplotlib.lines", "matplotlib.path",
    "matplotlib.gridspec", "matplotlib.colors", "matplotlib.transforms",
    "matplotlib.animation", "matplotlib.figure", "matplotlib.cm",
    "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d",
    "pypaperutils", "pypaperutils.design", "pypaperutils.io",
    "mypyutils", "mypyutils.misc", "mypyutils.io",
    "traci", "sumolib", "libsumo", "cv2",
]

_installed = False


def install():
    """Install stubs/substitutes and put the reference on ``sys.path``."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_SRC}")
    for name in _STUBS:
        if name in sys.modules:
            continue
        m = mock.MagicMock(name=name)
        m.__name__ = name
        m.__path__ = []
        m.__spec__ = None
        sys.modules[name] = m
    # classes used as base classes / in isinstance() must be real types
    sys.modules["matplotlib.axes"].Axes = type("Axes", (), {})

    def _read_yaml(path):
        import yaml
        with open(path, "r") as f:
            return yaml.safe_load(f)

    sys.modules["mypyutils.io"].read_yaml = _read_yaml
    sys.modules["mypyutils.misc"].none_switch = lambda a, b: b if a is None else a
    if "control" not in sys.modules:
        sys.modules["control"] = _build_control_module()
    for k, v in _build_bicycleparameters_modules().items():
        sys.modules.setdefault(k, v)
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)

    # constructor fix (reference vehicle.py:1359)
    from cyclistsocialforce import vehicle as rv
    from cyclistsocialforce.parameters import InvPendulumBicycleParameters
    from cyclistsocialforce.dynamics import PIDcontroller

    def _twod_init(self, s0, id="unknown", route=(), saveForces=False, params=None):
        if params is None:
            params = InvPendulumBicycleParameters()
        assert isinstance(params, InvPendulumBicycleParameters)
        rv.Bicycle.__init__(self, s0, id=id, route=route, saveForces=saveForces,
                            params=params)
        self.speed_controller = PIDcontroller(self.params.k_p_v, 0, 0,
                                              self.params.t_s, isangle=False)

    rv.TwoDBicycle.__init__ = _twod_init
    _installed = True


def modules():
    """Return (vehicle, intersection, parameters, dynamics, utils) reference modules."""
    install()
    from cyclistsocialforce import vehicle, intersection, parameters, dynamics, utils
    return vehicle, intersection, parameters, dynamics, utils


def headless_intersection(vehicles, **kw):
    """Reference SocialForceIntersection that steps without a matplotlib Axes
    (reference intersection.py:881-885 would call add_drawing(None))."""
    _, intersection, *_ = modules()
    ins = intersection.SocialForceIntersection(list(vehicles), **kw)
    ins.is_first_step = False
    return ins
