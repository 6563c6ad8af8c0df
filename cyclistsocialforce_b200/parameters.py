"""Parameter classes with the reference's names, defaults and error behaviour
(reference src/cyclistsocialforce/parameters.py).  Only the parameters the
stepping path consumes are kept; drawing parameters are out of scope.

They are the configuration surface of the engine: ``to_agent_params`` /
``to_field_params`` flatten them into the C structs the kernels take.
"""
from __future__ import annotations

import math

import numpy as np

from . import _lib
from . import whipplecarvallo as wc


def _check_float(name, v):
    if not isinstance(v, float):
        raise TypeError(f"{name} must be a float.")
    return v


def _check_pair(name, v):
    if not isinstance(v, (list, tuple, np.ndarray)) or len(v) != 2:
        raise TypeError(f"{name} must be a tuple of two floats (min, max).")
    if not v[0] <= v[1]:
        raise ValueError(f"{name}[0] must be <= {name}[1].")
    return [float(v[0]), float(v[1])]


class RoadElementParameters:
    """reference parameters.py:367-418 (F_0, sigma immutable, validated)."""

    def __init__(self, roadsurface_color=(0.8, 0.8, 0.8), roadedge_color="white",
                 roadedge_linewidth=1, F_0=0.05, sigma=3.0):
        self.roadsurface_color = roadsurface_color
        self.roadedge_color = roadedge_color
        self.roadedge_linewidth = roadedge_linewidth
        self.F_0 = F_0
        self.sigma = sigma

    @property
    def F_0(self):
        return self._F_0

    @F_0.setter
    def F_0(self, v):
        if hasattr(self, "_F_0"):
            raise AttributeError("F_0 is immutable.")
        _check_float("F_0", v)
        if not v >= 0:
            raise ValueError(f"F_0 must be >=0, instead it was {v:.2f}")
        self._F_0 = v

    @property
    def sigma(self):
        return self._sigma

    @sigma.setter
    def sigma(self, v):
        if hasattr(self, "_sigma"):
            raise AttributeError("sigma is immutable.")
        _check_float("sigma", v)
        if not v >= 0:
            raise ValueError(f"sigma must be >=0, instead it was {v:.2f}")
        self._sigma = v


class VehicleParameters:
    """Tactical + force-field parameters, reference parameters.py:421-750."""

    _FLOATS_POS = ("t_s", "d_arrived_inter", "d_arrived_stop", "v_max_stop", "v_max_harddecel")

    def __init__(self, t_s: float = 0.01, d_arrived_inter: float = 2.0, d_arrived_stop: float = 2.0,
                 v_max_stop: float = 0.1, v_max_harddecel: float = 2.5, hfov: float = 2 * np.pi,
                 calib_mode=False, verbose=True, rep_force=None, dest_force=None, dynamics=None,
                 f_0: float = 7.0, e_0: float = 0.995, e_1: float = 0.7, sigma_0: float = 0.5,
                 sigma_1: float = 5.0, sigma_2: float = 0.3, sigma_3: float = 4.9) -> None:
        self.calib_mode = calib_mode
        self.verbose = verbose
        self.rep_force = rep_force or {}
        self.dest_force = dest_force or {}
        self.dynamics = dynamics or {}
        for name, v in (("t_s", t_s), ("d_arrived_inter", d_arrived_inter),
                        ("d_arrived_stop", d_arrived_stop), ("v_max_stop", v_max_stop),
                        ("v_max_harddecel", v_max_harddecel)):
            _check_float(name, v)
            if not v >= 0:
                raise ValueError(f"{name} must be >=0, instead it was {v:.2f}")
            setattr(self, name, v)
        _check_float("hfov", float(hfov))
        if not 0 < hfov <= 2 * np.pi + 1e-12:
            raise ValueError(f"hfov must be in ]0,2pi], instead it was {hfov:.2f}")
        self.hfov = float(hfov)
        for name, v in (("f_0", f_0), ("e_0", e_0), ("e_1", e_1), ("sigma_0", sigma_0),
                        ("sigma_1", sigma_1), ("sigma_2", sigma_2), ("sigma_3", sigma_3)):
            _check_float(name, float(v))
            setattr(self, name, float(v))
        if not 0 <= self.e_0 < 1:
            raise ValueError("e_0 must be in [0,1[.")

    # -- flatten ---------------------------------------------------------------------------
    def to_field_params(self, q_scale: float, p2r: bool, field_kind: int = 0) -> "_lib.CsfFieldParams":
        fp = _lib.CsfFieldParams()
        fp.f_0, fp.e_0, fp.e_1 = self.f_0, self.e_0, self.e_1
        fp.sigma_0, fp.sigma_1, fp.sigma_2, fp.sigma_3 = self.sigma_0, self.sigma_1, self.sigma_2, self.sigma_3
        fp.hfov = self.hfov
        fp.q_scale = q_scale
        fp.p2r = 1 if p2r else 0
        fp.field_kind = int(field_kind)
        fp.p_0 = getattr(self, "p_0", 0.0)
        fp.p_decay = getattr(self, "p_decay", 1.0)
        fp.v_max = getattr(self, "v_max_riding", [0.0, 1.0])[1]
        # f32 tiled kernel: contributions below 2^-cutoff_log2 f_0 may be dropped (0 = library default 40;
        # set ``params.cutoff_log2 = 32`` to trade 18 % of the pair kernel's time for a looser bound)
        fp.cutoff_log2 = float(getattr(self, "cutoff_log2", 0.0))
        return fp

    def field_key(self, field_kind: int = 0):
        if field_kind == 1:
            return (1, self.p_0, self.p_decay, self.v_max_riding[1], self.hfov)
        return (0, self.f_0, self.e_0, self.e_1, self.sigma_0, self.sigma_1, self.sigma_2, self.sigma_3, self.hfov)

    def to_agent_params(self, q_scale: float, q_cap: int, hist_cap: int, origin=(0.0, 0.0)) -> "_lib.CsfAgentParams":
        p = _lib.CsfAgentParams()
        p.q_origin[0], p.q_origin[1] = float(origin[0]), float(origin[1])
        p.t_s = self.t_s
        p.d_arrived_inter, p.d_arrived_stop = self.d_arrived_inter, self.d_arrived_stop
        p.v_max_stop, p.v_max_harddecel = self.v_max_stop, self.v_max_harddecel
        for name, dst in (("a_max", p.a_max), ("a_desired_default", p.a_desired), ("v_max_riding", p.v_max_riding)):
            v = getattr(self, name, (0.0, 0.0))
            dst[0], dst[1] = float(v[0]), float(v[1])
        p.l = float(getattr(self, "l", 1.0))
        p.delta_max = float(getattr(self, "delta_max", 1.4))
        p.k_p_v = float(getattr(self, "k_p_v", 0.0))
        p.k_p_delta = float(getattr(self, "k_p_delta", 0.0))
        p.g = float(getattr(self, "g", 9.81))
        p.q_scale = q_scale
        p.traj_len = int(30 / self.t_s)          # vehicle.py:159
        p.hist_len = int(1 / self.t_s)           # vehicle.py:1487
        p.hist_cap = hist_cap
        p.q_cap = q_cap
        self._fill_model_params(p)
        return p

    def _fill_model_params(self, p):
        pass


class CarParameters(VehicleParameters):
    """reference parameters.py:753-763."""

    def __init__(self, length=4, width=2.0, **kwargs):
        super().__init__(**kwargs)
        self.length = length
        self.width = width


class BicycleParameters(VehicleParameters):
    """reference parameters.py:766-1173."""

    def __init__(self, v_max_riding: tuple = (-1.0, 10.0), v_desired_default: float = 5.0,
                 p_decay: float = 5.0, p_0: float = 30.0, hfov: float = np.pi * 2 / 3,
                 v_max_stop: float = 0.6, l: float = 1.0, l_1: float = None, l_2: float = None,
                 delta_max: float = 1.4, a_max: tuple = (-10.0, 10.0),
                 a_desired_default: tuple = (-5.0, 5.0), k_p_v: float = 10.0, k_p_delta: float = 10.0,
                 t_s: float = 0.01, d_arrived_inter: float = 2.0, d_arrived_stop: float = 2.0,
                 v_max_harddecel: float = 2.5, g=9.81, **kwargs) -> None:
        super().__init__(t_s=t_s, d_arrived_inter=d_arrived_inter, d_arrived_stop=d_arrived_stop,
                         v_max_stop=v_max_stop, v_max_harddecel=v_max_harddecel, hfov=hfov, **kwargs)
        self.v_max_riding = _check_pair("v_max_riding", v_max_riding)
        self.a_max = _check_pair("a_max", a_max)
        self.a_desired_default = _check_pair("a_desired_default", a_desired_default)
        self.v_desired_default = v_desired_default
        self.p_decay = float(p_decay)
        self.p_0 = float(p_0)
        # wheelbase bookkeeping (reference :930-990): l = l_1 + l_2
        if l is None:
            l_1 = 0.5 if l_1 is None else l_1
            l_2 = 0.5 if l_2 is None else l_2
            l = l_1 + l_2
        else:
            if l_1 is None and l_2 is None:
                l_1 = l_2 = l / 2
            elif l_1 is None:
                l_1 = l - l_2
            elif l_2 is None:
                l_2 = l - l_1
        self.l, self.l_1, self.l_2 = float(l), float(l_1), float(l_2)
        self.delta_max = _check_float("delta_max", float(delta_max))
        self.k_p_v = _check_float("k_p_v", float(k_p_v))
        self.k_p_delta = _check_float("k_p_delta", float(k_p_delta))
        self.g = float(g)

    @property
    def v_desired_default(self):
        return self._v_desired_default

    @v_desired_default.setter
    def v_desired_default(self, v):
        if not isinstance(v, (float, int)):
            raise TypeError("v_desired_default must be a float.")
        self._v_desired_default = float(v)


class InvPendulumBicycleParameters(BicycleParameters):
    """reference parameters.py:1414-1970 (defaults :1429-1472)."""

    #: full-state feedback gain polynomials in 1/v, reference parameters.py:1863-1883
    KX_TABLE = (
        (3.48203226e02, -5.12057324e03, 1.58364873e04, -1.98073306e04),
        (-4.51700000e01, 0.0, 0.0, 0.0),
        (-9.16379250e02, 1.31769807e04, -6.57341643e04, 8.22163589e04),
        (3.20214069e02, -4.69953797e03, 1.66378680e04, -2.43114309e04),
        (2.87549256e-08, -2.27913445e03, 0.0, 0.0),
    )
    KU_TABLE = (-3.38638984e-09, -2.27913445e03, 0.0, 0.0)

    def __init__(self, v_max_riding: tuple = (-1.0, 7.0), v_desired_default: float = 5.0,
                 hfov: float = np.pi * 2 / 3, a_max: tuple = (-3.0, 1.0),
                 a_desired_default: tuple = (-1.0, 0.5), l: float = None, l_1: float = 0.5,
                 l_2: float = 0.5, delta_max: float = 1.4, h: float = 1.0, m: float = 87.0,
                 i_bike_longlong: float = 3.28, i_steer_vertvert: float = 0.07, c_steer: float = 50.0,
                 k_p_v: float = 10.0, k_d0_r2: float = -600.0, k_d1_r2: float = 0.2,
                 k_p_r1: float = 0.25, k_i0_r1: float = 0.2, t_s: float = 0.01,
                 d_arrived_inter: float = 2.0, d_arrived_stop: float = 2.0, v_max_harddecel: float = 2.5,
                 v_max_stop: float = 0.6, v_max_walk: float = 1.5, delta_max_walk: float = 0.174,
                 f_0: float = 7.0, e_0: float = 0.995, e_1: float = 0.7, sigma_0: float = 0.5,
                 sigma_1: float = 5.0, sigma_2: float = 0.3, sigma_3: float = 4.9, g: float = 9.81) -> None:
        super().__init__(v_max_riding=v_max_riding, v_desired_default=v_desired_default, hfov=hfov,
                         a_max=a_max, a_desired_default=a_desired_default, l=l, l_1=l_1, l_2=l_2,
                         delta_max=delta_max, k_p_v=k_p_v, t_s=t_s, d_arrived_inter=d_arrived_inter,
                         d_arrived_stop=d_arrived_stop, v_max_harddecel=v_max_harddecel,
                         v_max_stop=v_max_stop, g=g, f_0=f_0, e_0=e_0, e_1=e_1, sigma_0=sigma_0,
                         sigma_1=sigma_1, sigma_2=sigma_2, sigma_3=sigma_3)
        for name, v in (("h", h), ("m", m), ("i_bike_longlong", i_bike_longlong),
                        ("i_steer_vertvert", i_steer_vertvert), ("c_steer", c_steer),
                        ("v_max_walk", v_max_walk), ("delta_max_walk", delta_max_walk)):
            _check_float(name, v)
            setattr(self, name, v)
        self.k_d0_r2, self.k_d1_r2, self.k_p_r1, self.k_i0_r1 = k_d0_r2, k_d1_r2, k_p_r1, k_i0_r1

    @property
    def tau_1_squared(self):
        return (self.i_bike_longlong + self.m * self.h ** 2) / (self.m * self.g * self.h)

    def timevarying_combined_params(self, v: float):
        """reference parameters.py:1832-1855."""
        return (v ** 2) / (self.g * self.l), (v * self.l_2) / (self.g * self.l), self.l / v

    def fullstate_feedback_gains(self, v):
        """reference parameters.py:1857-1892."""
        vdata = np.array((1, v ** -1, v ** -2, v ** -3))
        return (np.array(self.KX_TABLE) @ vdata)[np.newaxis, :], np.array(self.KU_TABLE) @ vdata

    def _fill_model_params(self, p):
        p.l_2 = self.l_2
        p.tau_1_squared = self.tau_1_squared
        p.i_steer = self.i_steer_vertvert
        p.c_steer = self.c_steer
        p.v_max_walk = self.v_max_walk
        p.delta_max_walk = self.delta_max_walk
        for r in range(5):
            for c in range(4):
                p.kx_table[r][c] = self.KX_TABLE[r][c]
        for c in range(4):
            p.ku_table[c] = self.KU_TABLE[c]


class PlanarPointBicycleParameters(BicycleParameters):
    """reference parameters.py:1175-1201."""

    def __init__(self, poles=(-2 + 0j,), gains=(2,), **kwargs):
        super().__init__(**kwargs)
        self.gains = gains
        self.poles = [(-2 + 0j)] if poles is None else [poles[0]]

    def _fill_model_params(self, p):
        # PlanarPointDynamics._get_gains, reference dynamics.py:933-940 (poles win over gains)
        p.k_psi = float(-np.real(self.poles[0])) if self.poles is not None else float(self.gains[0])


class BalancingRiderBicycleParameters(BicycleParameters):
    """reference parameters.py:1214-1411.

    The rider-behaviour ("pole") models ship as the linear pole-vs-speed regressions the
    reference derives at construction (``PoleModel.get_component_mean_function``,
    controlbehavior.py:1601-1650); they were extracted with tests/golden/make_golden.py.
    ``stochastic_control_behavior=True``: every rider draws its closed-loop poles from the model -- a
    Gaussian mixture over [speed, pole features] conditioned on the speed -- and re-draws them whenever its
    speed has moved by more than ``controlparam_resampling_speedthresh`` (reference :1398-1402,
    controlbehavior.py:1414-1469).  The draws happen on the device with a counter-based generator
    (``controlparam_seed``; csrc/csf_agent.cu ``br_sample_poles``); the model's numbers ship as
    data/pole_models.json (tests/golden/make_polemodels.py re-serialises the reference's YAML files).
    """

    #: features [p0_real, p1_real, p1_imag, p2_real, p2_imag]: (intercept, slope) per model/component
    POLE_REGRESSIONS = {
        ("BR1_ImRe5GivenV_pole-model-params.yaml", 0): (
            (-1.8535775147013691, -0.17961038648194527, 1.1730293099635356, 0.6329343335691049,
             2.3198026534882037),
            (-1.6314600197495845, -0.15971107558265726, 0.14249117463843872, -0.44818335654213604,
             0.9166444997604464)),
        ("BR1_ImRe5GivenV_pole-model-params.yaml", 1): (
            (-0.35234247428779364, -0.10273325648260068, 1.738236648818369, -0.47097770394470295,
             7.811765217705167),
            (-0.588378287364706, -0.28745435083339654, 0.3111051504915779, -0.7008198641061155,
             0.06086016583769994)),
        ("BR0_ImRe5GivenV_pole-model-params.yaml", 0): (
            (7.477367764370239, -0.6066675056522426, 1.7881981548329726, -1.32823781934098,
             5.327111219864689),
            (-7.589580229524327, -0.10886032204606368, 0.04106272039749894, -0.02666449568231668,
             0.08910709292445088)),
    }

    def __init__(self, bicycleParameterDict=None, poles=None, gains=None,
                 controlparam_filename="BR1_ImRe5GivenV_pole-model-params.yaml",
                 stochastic_control_behavior=False, controlparam_resampling_speedthresh=0.8333,
                 controlparam_polemodel_component=0, p_dist_roll=0.00, p_dist_steer=0.00,
                 T_dist_roll=9000, T_dist_steer=1000, controlparam_seed=0, **kwargs):
        if bicycleParameterDict is None:
            bicycleParameterDict = wc.balanceassistv1_with_averagerider
        self.bike = dict(bicycleParameterDict)
        kwargs = dict(kwargs, l=self.bike["w"], l_1=self.bike["w"] / 2)      # reference :1290-1292
        super().__init__(**kwargs)
        self.m = self.bike["mB"] + self.bike["mF"] + self.bike["mH"] + self.bike["mR"]
        self.g = self.bike["g"]
        if p_dist_roll > 0 or p_dist_steer:
            raise Warning("Support for steer and roll torque disturbance removed!")  # dynamics.py:317-318
        # fixed control parameters (reference :1306-1314): the pole model is ignored; ``gains`` wins over
        # ``poles`` (dynamics.py:606-607 returns params.gains before it ever looks at the poles)
        self.controlparam_fix = poles is not None or gains is not None
        self.poles = None if poles is None else np.asarray(poles, dtype=complex).reshape(-1)
        self.gains = None if gains is None else np.asarray(gains, dtype=float).reshape(-1)
        if self.gains is not None and self.gains.shape != (5,):
            raise ValueError("gains: the five full-state feedback gains K_x (the last one is also the input gain)")
        fixed_features = None
        if self.poles is not None:
            fixed_features = _pole_features(self.poles)
        if self.controlparam_fix:
            stochastic_control_behavior = False
        self.stochastic_control_behavior = bool(stochastic_control_behavior)
        self.controlparam_filename = controlparam_filename
        self.controlparam_polemodel_component = controlparam_polemodel_component
        self.controlparam_resampling_speedthresh = float(controlparam_resampling_speedthresh)
        self.controlparam_seed = int(controlparam_seed)
        self.polemodel = None
        if self.stochastic_control_behavior:
            models = load_pole_models()
            if controlparam_filename not in models:
                raise FileNotFoundError(
                    f"Couldn't find Balancing Rider Control Behavior model {controlparam_filename}. "
                    f"Available models are: {sorted(models)}")
            self.polemodel = models[controlparam_filename]
            n_comp = len(self.polemodel["weights"])
            if controlparam_polemodel_component >= n_comp:                    # reference :1369-1373
                raise ValueError(f"Balancing Rider Control Behavior model {controlparam_filename} has only {n_comp} "
                                 f"components but controlparam_polemodel_component is set to "
                                 f"{controlparam_polemodel_component}!")
            if self.polemodel["index_given"] != 0 or len(self.polemodel["features"]) != 6 or n_comp > 4:
                raise NotImplementedError("pole model layout not supported by the device sampler")
        key = (controlparam_filename, controlparam_polemodel_component if not self.stochastic_control_behavior else 0)
        if self.controlparam_fix:
            f = fixed_features if fixed_features is not None else (-1.0, -2.0, 1.0, -3.0, 1.0)   # (unused with gains)
            self.pole_intercept, self.pole_slope = tuple(f), (0.0,) * 5
        else:
            if key not in self.POLE_REGRESSIONS:
                raise FileNotFoundError(
                    f"Couldn't find Balancing Rider Control Behavior model {key}. "
                    f"Available models are: {sorted(self.POLE_REGRESSIONS)}")
            self.pole_intercept, self.pole_slope = self.POLE_REGRESSIONS[key]
        self.p_dist_roll, self.p_dist_steer = p_dist_roll, p_dist_steer
        self.T_dist_roll, self.T_dist_steer = T_dist_roll, T_dist_steer

    def poles_at(self, v):
        """update_control_params, reference parameters.py:1403-1411."""
        f = np.asarray(self.pole_intercept) + np.asarray(self.pole_slope) * v
        return [f[0] + 0j, f[1] + 1j * f[2], f[1] - 1j * f[2], f[3] + 1j * f[4], f[3] - 1j * f[4]]

    def _fill_model_params(self, p):
        A0, A1, A2, B = wc.speed_polynomial_state_matrices(self.bike)
        for i in range(25):
            p.br_A0[i], p.br_A1[i], p.br_A2[i] = A0.flat[i], A1.flat[i], A2.flat[i]
        for i in range(5):
            p.br_B[i] = B[i]
            p.br_pole_icpt[i] = self.pole_intercept[i]
            p.br_pole_coef[i] = self.pole_slope[i]
        p.br_stochastic = 1 if self.stochastic_control_behavior else 0
        p.br_fixed_gains = 1 if self.gains is not None else 0
        if self.stochastic_control_behavior:
            m = self.polemodel
            cov, mu, w = np.array(m["covariances"]), np.array(m["means"]), np.array(m["weights"])
            p.br_n_comp = len(w)
            p.br_resample_thresh = self.controlparam_resampling_speedthresh
            p.br_seed = self.controlparam_seed & 0xFFFFFFFFFFFFFFFF
            for i in range(6):
                p.br_lam[i], p.br_sc_mean[i], p.br_sc_scale[i] = m["lambdas"][i], m["scaler_mean"][i], m["scaler_scale"][i]
            logf = list(m["log_features"])
            for i in range(5):                       # pole feature i = model feature i + 1
                j = logf.index(i + 1) if (i + 1) in logf else -1
                p.br_log_a[i] = m["log_a"][j] if j >= 0 else 0.0
                p.br_log_sign[i] = m["log_sign"][j] if j >= 0 else 0.0
            for c in range(len(w)):
                # Gaussian component conditioned on the speed (feature 0): mean moves linearly with it, the
                # covariance is the Schur complement (controlbehavior.py:477-533)
                var_g, cg = cov[c, 0, 0], cov[c, 1:, 0]
                L = np.linalg.cholesky(cov[c, 1:, 1:] - np.outer(cg, cg) / var_g)
                p.br_w[c], p.br_mu_g[c], p.br_var_g[c] = w[c], mu[c, 0], var_g
                for i in range(5):
                    p.br_mu[c][i] = mu[c, 1 + i]
                    p.br_slope[c][i] = cg[i] / var_g
                    for j in range(i + 1):
                        p.br_chol[c][i * (i + 1) // 2 + j] = L[i, j]


def _pole_features(poles):
    """Five closed-loop poles (one real, two complex-conjugate pairs, any order) -> the features
    [p0_real, p1_real, p1_imag, p2_real, p2_imag] the gain design works with."""
    poles = np.asarray(poles, dtype=complex).reshape(-1)
    if poles.shape != (5,):
        raise ValueError("poles: five closed-loop poles (one real, two complex-conjugate pairs)")
    real = [z for z in poles if abs(z.imag) < 1e-14]
    upper = sorted([z for z in poles if z.imag >= 1e-14], key=lambda z: (z.real, z.imag), reverse=True)
    lower = [z for z in poles if z.imag <= -1e-14]
    if len(real) == 5 or len(real) == 3:
        # real pairs: (A - a)(A - b) = A^2 - (a + b) A + a b  ==  re = (a + b)/2, im^2 = a b - re^2 < 0 is not
        # representable with a real imaginary part
        raise NotImplementedError("fixed poles: one real pole and two complex-conjugate pairs are supported")
    if len(real) != 1 or len(upper) != 2 or len(lower) != 2:
        raise ValueError("poles must be one real pole and two complex-conjugate pairs")
    for z in upper:
        if min(abs(z.conjugate() - w) for w in lower) > 1e-9 * max(1.0, abs(z)):
            raise ValueError("poles must come in complex-conjugate pairs")
    return (float(real[0].real), float(upper[0].real), float(upper[0].imag), float(upper[1].real), float(upper[1].imag))


_POLE_MODELS = None


def load_pole_models():
    """The rider-behaviour models packaged with the reference (data/balancingriderparams/*.yaml), as
    re-serialised numbers: data/pole_models.json."""
    global _POLE_MODELS
    if _POLE_MODELS is None:
        import json
        import os
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "pole_models.json")) as f:
            _POLE_MODELS = json.load(f)
    return _POLE_MODELS


def payload_frame(points, margin_rel=0.25, margin_abs=250.0):
    """Origin and half-extent of the Q-format frame for a crowd: the centre of the bounding box of
    ``points`` (iterable of (..., 2) arrays: positions, destinations, obstacles) and its half-size plus a
    margin for road users that leave the box.  A road user outside origin +- extent raises status bit 2."""
    lo = np.array([np.inf, np.inf])
    hi = -lo
    for a in points:
        a = np.asarray(a, float).reshape(-1, 2)
        if a.size:
            lo, hi = np.minimum(lo, a.min(axis=0)), np.maximum(hi, a.max(axis=0))
    if not np.all(np.isfinite(lo)):
        return (0.0, 0.0), 1000.0
    half = float((hi - lo).max()) / 2
    return (float((lo[0] + hi[0]) / 2), float((lo[1] + hi[1]) / 2)), half * (1.0 + margin_rel) + margin_abs


def choose_q_scale(extent_m: float) -> float:
    """Metres per unit of the Q-format int32 payload: the finest power of two such that
    +-extent fits in +-2^30 units."""
    extent_m = max(float(extent_m), 1.0)
    k = math.floor(math.log2((2.0 ** 30) / extent_m))
    return 2.0 ** (-k)
