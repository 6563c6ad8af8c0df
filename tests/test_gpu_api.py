"""The reference-facing Python API on the GPU: scenarios written the way the reference's
demo / scenario scripts write them, checked against the oracle and the golden vectors."""
import numpy as np
import pytest
import torch

from oracle import csf_oracle as co
from helpers import oracle_world
from cyclistsocialforce_b200 import parameters as P
from cyclistsocialforce_b200.intersection import (CurvedRoadSegment, RoadSegmentCollection,
                                                  SocialForceIntersection, StraightRoadSegment)
from cyclistsocialforce_b200.scenario import Scenario
from cyclistsocialforce_b200.vehicle import (BalancingRiderBicycle, InvPendulumBicycle, PlanarPointBicycle,
                                             TwoDBicycle, UncontrolledVehicle)

pytestmark = pytest.mark.gpu


def _demo_bikes(cls):
    """demo/demoCSFstandalone.py:101-118."""
    bike1 = cls((-23 + 17, 0, 0, 5, 0, 0, 0, 0), id="a", saveForces=True)
    bike1.params.v_desired_default = 4.5
    bike2 = cls((0 + 15, -20, np.pi / 2, 5, 0, 0, 0, 0), id="b", saveForces=True)
    bike2.params.v_desired_default = 5.0
    bike3 = cls((-2 + 15, -20, np.pi / 2, 5, 0, 0, 0, 0), id="c", saveForces=True)
    bike3.params.v_desired_default = 5.0
    bike1.setDestinations((35, 64, 65), (0, 0, 0))
    bike2.setDestinations((15, 15, 15), (20, 49, 50))
    bike3.setDestinations((13, 13, 13), (20, 49, 50))
    return bike1, bike2, bike3


@pytest.mark.parametrize("cls,model", [(TwoDBicycle, "twod"), (PlanarPointBicycle, "planarpoint"),
                                       (InvPendulumBicycle, "invpendulum")])
def test_demo_scenario_f64(golden, cls, model):
    """BASELINE config 1, written like the reference's DemoScenario."""
    bikes = _demo_bikes(cls)

    class DemoScenario(Scenario):
        def __init__(self):
            self.intersection = SocialForceIntersection(bikes, activate_sumo_cosimulation=False,
                                                        dtype=torch.float64)
            Scenario.__init__(self, self._step_func, verbose=False, realtime=False)

        def _step_func(self):
            self.intersection.step()

    scn = DemoScenario()
    scn.run(7)
    assert scn.i == 700
    s = np.array([b.s for b in scn.intersection.vehicles])
    k = golden[f"demo_{model}_steps"].tolist().index(700)
    assert np.abs(s - golden[f"demo_{model}_s"][k]).max() < 1e-8
    f = np.array([b.force for b in scn.intersection.vehicles])
    assert np.abs(f - golden[f"demo_{model}_F"][k]).max() < 1e-8
    # histories like the reference keeps them
    b = scn.intersection.vehicles[0]
    assert b.i == 700 and len(b.F) == 700 and np.array_equal(b.traj[:, 700], b.s)
    assert np.allclose(b.trajF[:, 700], f[0])
    assert scn.intersection.hist_n_vecs == [3] * 700


def test_demo_scenario_f32_drift(golden):
    """fp32 production build on the demo: bounded trajectory drift over 700 steps (reported)."""
    ins = SocialForceIntersection(_demo_bikes(TwoDBicycle), dtype=torch.float32)
    for _ in range(700):
        ins.step()
    s = np.array([b.s for b in ins.vehicles])
    ref = golden["demo_twod_s"][golden["demo_twod_steps"].tolist().index(700)]
    drift = np.abs(s - ref)
    from test_gpu_parity import report
    report(test="demo_f32_drift_700_steps", max_pos_m=float(drift[:, :2].max()), max_other=float(drift[:, 2:].max()))
    assert drift[:, :2].max() < 5e-3 and drift[:, 2:].max() < 5e-3


def test_parcours_scenario_f64(golden):
    """BASELINE config 2 (scenarios/parcours-scenario.py:31-40)."""
    ins = SocialForceIntersection([], dtype=torch.float64)
    b = BalancingRiderBicycle((0, 0, np.pi / 2, 5, 0, 0, 0, 0), id="BalancingRiderBike", saveForces=True)
    b.params.v_desired_default = 4.0
    destx = [0, 10, 0, 5, 10, 20, 21, 22, 23]
    desty = [10, 20, 30, 40, 40, 40, 40, 40, 40]
    b.setDestinations(destx, desty)
    ins.add_road_user(b)
    scn = Scenario(ins.step, verbose=False, realtime=False)
    scn.run(15)
    assert np.abs(b.s - golden["parcours_s"][-1][0]).max() < 1e-7
    assert b.destpointer == int(golden["parcours_ptr"][0])


def test_add_and_remove_road_users():
    """Churn (reference :458-539, :576-634): results equal a crowd that never changed."""
    s0, q = co.synthetic_crowd(12, seed=4, spacing=3.0)

    def mk(k):
        b = TwoDBicycle(tuple(s0[k]), id=f"v{k}")
        b.setDestinations(q[k, :, 0], q[k, :, 1])
        return b

    ins = SocialForceIntersection([mk(k) for k in range(12)], dtype=torch.float64)
    ref = SocialForceIntersection([mk(k) for k in range(12)], dtype=torch.float64)
    for _ in range(30):
        ins.step()
        ref.step()
    extra = TwoDBicycle((200.0, 200.0, 0.0, 5.0, 0.0), id="far")      # far away: negligible interaction
    extra.setDestinations((260.0,), (200.0,))
    ins.add_road_user(extra)
    assert ins.n_bikes == 13 and ins.get_road_user_ids()[-1] == "far"
    for _ in range(5):
        ins.step()
        ref.step()
    ins.remove_road_users_by_id(["far"])
    assert ins.n_bikes == 12 and extra._owner is None
    for _ in range(20):
        ins.step()
        ref.step()
    a = np.array([v.s for v in ins.vehicles])
    b = np.array([v.s for v in ref.vehicles])
    assert np.abs(a - b).max() < 1e-9
    assert [v.i for v in ins.vehicles] == [55] * 12


@pytest.mark.parametrize("cls_name,model", [("TwoDBicycle", "twod"), ("InvPendulumBicycle", "invpendulum")])
def test_churn_against_the_oracle(cls_name, model):
    """add_road_user / remove_road_user / remove_road_users_by_id (reference intersection.py:458-539,
    :576-634) in the middle of a run, against the ORACLE doing the same list operations on per-vehicle
    state (the reference's vehicles own their state): the road users that stay keep their navigation
    machine, history and dynamic state; a road user that joins starts from its constructor state."""
    import cyclistsocialforce_b200.vehicle as V
    cls = getattr(V, cls_name)
    ns = cls.N_STATES
    n = 14
    s0, q = co.synthetic_crowd(n + 2, seed=6, spacing=3.0, n_states=ns)

    def mk(k):
        b = cls(tuple(s0[k]), id=f"v{k}")
        b.setDestinations(q[k, :, 0], q[k, :, 1])
        return b

    def mk_oracle(k):
        A = co.Agents(model, s0[k:k + 1])
        A.set_destinations(0, q[k, :, 0], q[k, :, 1])
        A.tag = f"v{k}"
        return A

    ins = SocialForceIntersection([mk(k) for k in range(n)], dtype=torch.float64)
    W = co.World([mk_oracle(k) for k in range(n)])

    def check(tag):
        got = {v.id: v.s for v in ins.vehicles}
        assert [v.id for v in ins.vehicles] == [g.tag for g in W.groups], tag
        for g in W.groups:
            assert np.abs(got[g.tag] - g.s[0]).max() < 1e-8, (tag, g.tag, np.abs(got[g.tag] - g.s[0]).max())

    def run(k):
        for _ in range(k):
            ins.step()
            W.step()

    run(12)
    check("before churn")
    pool = {k: v.data_ptr() for k, v in ins._engine._pool.items()}
    ins.remove_road_user(3)                                  # reference :576-616 pops index 3
    W.groups.pop(3)
    # the replaced engine handed its device buffers over: no allocation for a crowd that did not grow
    assert pool and all(ins._engine._pool[k].data_ptr() == ptr for k, ptr in pool.items())
    run(9)
    check("after remove_road_user")
    ins.add_road_user(mk(n))                                 # a fresh road user in the middle of the crowd
    W.groups.append(mk_oracle(n))
    run(11)
    check("after add_road_user")
    ins.remove_road_users_by_id(["v0", "v7", f"v{n}"])       # :618-634
    W.groups = [g for g in W.groups if g.tag not in ("v0", "v7", f"v{n}")]
    ins.add_road_user(mk(n + 1))
    W.groups.append(mk_oracle(n + 1))
    run(15)
    check("after remove_road_users_by_id + add")
    assert ins.n_bikes == n - 2 and ins.hist_n_vecs[-1] == n - 2
    ins.check_status()


def test_vehicle_hooks_match_oracle():
    """Vehicle.calcRepulsiveForce / calcDestinationForce / step used on their own."""
    b = TwoDBicycle((1.0, 2.0, 0.4, 5.0, 0.0), id="solo")
    b.setDestinations((40.0, 80.0), (10.0, 30.0))
    ins = SocialForceIntersection([b], dtype=torch.float64)
    rng = np.random.default_rng(0)
    x, y, psi = rng.uniform(-8, 10, 50), rng.uniform(-8, 10, 50), rng.uniform(-3, 3, 50)
    fx, fy = b.calcRepulsiveForce(x, y, psi)
    p = co.default_params("twod")
    ox, oy = co.twod_field(1.0, 2.0, 0.4, co.field_params_array([p])[0], x, y, psi)
    assert np.abs(fx - ox).max() < 1e-12 and np.abs(fy - oy).max() < 1e-12
    A = co.Agents("twod", np.array([[1.0, 2.0, 0.4, 5.0, 0.0]]))
    A.set_destinations(0, (40.0, 80.0), (10.0, 30.0))
    for _ in range(5):
        f = b.calcDestinationForce()
        fo = A.calc_destination_force(0)
        assert np.abs(np.array(f) - np.array(fo)).max() < 1e-11
        b.step(*f)
        A.step_agent(0, *fo)
        assert np.abs(b.s - A.s[0]).max() < 1e-12
    assert b.i == 5


def test_obstacle_and_mixed_parameters():
    """UncontrolledVehicle sources (hfov = 2 pi) + two bicycle parameter sets in one intersection."""
    s0, q = co.synthetic_crowd(10, seed=8, spacing=2.5)
    wide = P.InvPendulumBicycleParameters(hfov=np.pi, f_0=5.0)
    bikes = []
    for k in range(10):
        b = TwoDBicycle(tuple(s0[k]), id=f"v{k}", params=wide if k >= 6 else None)
        b.setDestinations(q[k, :, 0], q[k, :, 1])
        bikes.append(b)
    traj = np.array([[3.0 + 0.02 * t, 3.0, 0.0, 2.0] for t in range(200)]).T
    car = UncontrolledVehicle((3.0, 3.0, 0.0, 2.0), trajectory=traj, id="car")
    ins = SocialForceIntersection(bikes + [car], dtype=torch.float64)
    g1 = co.Agents("twod", s0[:6], destqueue=None)
    g2 = co.Agents("twod", s0[6:], params=co.default_params("twod", hfov=np.pi, f_0=5.0))
    for k in range(6):
        g1.set_destinations(k, q[k, :, 0], q[k, :, 1])
    for k in range(4):
        g2.set_destinations(k, q[6 + k, :, 0], q[6 + k, :, 1])
    g3 = co.Agents("uncontrolled", np.array([[3.0, 3.0, 0.0, 2.0]]), uncontrolled_traj=[traj])
    W = co.World([g1, g2, g3])
    for _ in range(40):
        ins.step()
        W.step()
    got = np.array([b.s for b in bikes])
    ref = np.vstack([g1.s, g2.s])
    assert np.abs(got - ref).max() < 1e-9
    assert abs(car.s[0] - g3.s[0, 0]) < 1e-15


def test_road_elements_in_intersection():
    """scenarios/curve-scenario.py geometry: road-edge force added after the clip (:854-857)."""
    roadparams = P.RoadElementParameters(sigma=2.0, F_0=0.15)
    seg1 = StraightRoadSegment(np.array((0, -20, np.pi / 2)), 5, 25, params=roadparams, ds=0.1)
    seg2 = CurvedRoadSegment(seg1.x1, 5, 10, np.pi / 2, "right", params=roadparams, ds=0.1)
    segs = RoadSegmentCollection((seg1, seg2))
    bikes = []
    for k, (x, y) in enumerate([(0.5, -19.0), (-0.8, -15.0)]):
        b = TwoDBicycle((x, y, np.pi / 2, 4.0, 0.0), id=f"r{k}")
        b.setDestinations((0.0, 3.0, 10.0), (0.0, 12.0, 15.0))
        bikes.append(b)
    ins = SocialForceIntersection(bikes, road_elements=[segs], dtype=torch.float64)
    A = co.Agents("twod", np.array([b.s for b in bikes]))
    for k in range(2):
        A.set_destinations(k, (0.0, 3.0, 10.0), (0.0, 12.0, 15.0))
    W = co.World([A], road_edges=segs.edges_flat())
    for _ in range(150):
        ins.step()
        W.step()
    assert np.abs(np.array([b.s for b in bikes]) - A.s).max() < 1e-9
    fx, fy = ins.calc_forces()
    ofx, ofy = W.calc_forces()
    assert np.abs(fx - ofx).max() < 1e-9 and np.abs(fy - ofy).max() < 1e-9


F32_STEP_CASES = dict(twod=(768, 16), planarpoint=(320, 10), invpendulum=(320, 10), balancingrider=(128, 8),
                      bicycle=(384, 12))


@pytest.mark.parametrize("model", list(F32_STEP_CASES))
def test_f32_per_step_error(model):
    """north_star: the fp32 production build stays within 1e-4 relative PER STEP of the reference
    arithmetic.  Three crowds step side by side -- the oracle, the f64 build and the f32 build -- and
    before every step the f32 build is reset to the f64 build's state (every per-agent field), so each
    f32 step starts from the state the oracle step starts from (the f64 build tracks the oracle to
    1e-9, asserted here).  The f32 forces and states are compared with the ORACLE's, every agent, every
    step.  The only agents exempt are those with a source within 1e-5 rad of their field-of-view
    boundary at the start of the step (the mask is discontinuous there: the reference's own result
    flips under a 1e-16 perturbation); their number is counted, bounded and reported."""
    from gpu_helpers import make_engine
    from helpers import oracle_world
    from test_gpu_parity import report, _vec_rel, _rel
    from cyclistsocialforce_b200.engine import _FIELD_AXIS, _STATE_COLS
    n, steps = F32_STEP_CASES[model]
    s0, q = co.synthetic_crowd(n, seed=33, spacing=3.0, n_states=8)
    W = oracle_world(model, s0, np.full(n, 5.0), q)
    A = W.groups[0]
    e64, g64 = make_engine(model, s0, 5.0, q, dtype=torch.float64)
    e32, g32 = make_engine(model, s0, 5.0, q, dtype=torch.float32, q_scale=e64.q_scale, origin=e64.q_origin)
    fields = [f for f in _STATE_COLS + tuple(_FIELD_AXIS) if f not in ("destq", "dest_len", "vd_default")
              and getattr(g64, f) is not None]
    fp = co.field_params_array([A.p])[0]
    ns = A.s.shape[1]
    lin = [0, 1] + list(range(3, ns))                       # x, y, v (+ rates): relative to max(|.|, 1)
    ang = [2] + [c for c in (4, 5) if c < ns]               # psi, delta, theta: absolute, rad
    lin = [c for c in lin if c not in ang]
    worst = dict(f32_force=0.0, f32_state=0.0, f64_force=0.0, f64_state=0.0)
    exempt = exempt_steer = 0
    for step in range(steps):
        for name in fields:
            getattr(g32, name).copy_(getattr(g64, name).to(getattr(g32, name).dtype))
        e32.pack()
        _, margin = co.pair_forces(A.s[:, 0], A.s[:, 1], A.s[:, 2], fp, return_margin=True)
        W.step()
        e64.step()
        e32.step()
        fo, so = A.force, A.s
        f64, s64 = e64.force.cpu().numpy(), g64.states_numpy()
        f32, s32 = e32.force.cpu().numpy().astype(float), g32.states_numpy()

        # the total force is a sum of two terms of magnitude |F_dest| that may cancel: its error is judged
        # against the terms (|F_rep| <= |F_dest| after the clip, intersection.py:841-848), at least 1
        fscale = np.maximum(np.maximum(np.linalg.norm(fo, axis=1), np.linalg.norm(A.fdest, axis=1)), 1.0)

        def errs(f, s):
            ef = np.linalg.norm(f - fo, axis=1) / fscale
            es = _rel(s[:, lin], so[:, lin], 1.0).max(axis=1)
            for c in ang:
                es = np.maximum(es, np.abs(np.angle(np.exp(1j * (s[:, c] - so[:, c])))))
            return ef, es

        ef64, es64 = errs(f64, s64)
        worst["f64_force"] = max(worst["f64_force"], ef64.max())
        worst["f64_state"] = max(worst["f64_state"], es64.max())
        assert ef64.max() < 1e-8 and es64.max() < 1e-8, (model, step, ef64.max(), es64.max())
        ef, es = errs(f32, s32)
        ok = margin > 1e-5
        state_tol = 1e-4 + ef * fscale / np.maximum(np.linalg.norm(fo, axis=1), 1e-300)
        exempt += int((~ok).sum())
        exempt_steer += int((ok & (es > 1e-4)).sum())          # agent-steps that needed the |dF| / |F| term
        worst["f32_force"] = max(worst["f32_force"], ef[ok].max())
        worst["f32_state"] = max(worst["f32_state"], (es / state_tol * 1e-4)[ok].max())
        bad = ok & ((ef > 1e-4) | (es > state_tol))
        assert not bad.any(), (model, step, np.flatnonzero(bad)[:8], ef[bad][:8], es[bad][:8], margin[bad][:8],
                               np.linalg.norm(fo, axis=1)[bad][:8])
    e32.check_status()
    e64.check_status()
    report(test="f32_per_step_vs_oracle", model=model, n=n, steps=steps, agent_steps=n * steps,
           exempt_fov_boundary=exempt, state_over_1e4_by_cancelled_force=exempt_steer,
           **{k: float(v) for k, v in worst.items()})
    # a source within 1e-5 rad of the boundary of a 2.09 rad cone: ~n * 2 * 1e-5 / (2 pi) per target and step
    assert exempt <= max(3, int(20 * n * steps * n * 2e-5 / (2 * np.pi)))
    assert exempt_steer <= n * steps // 50


def test_graph_step_equals_kernel_by_kernel_step():
    """Engine(graph=True) replays a CUDA graph of the step; the periodic re-sort of the spatial order
    stays outside the graph.  Same kernels, same order -> bitwise the same crowd, across re-sorts."""
    from cyclistsocialforce_b200 import parameters as P
    from cyclistsocialforce_b200.engine import AgentGroup, Engine
    from cyclistsocialforce_b200.synthetic import queues_with_start
    n = 6000
    s0, q = co.synthetic_crowd(n, seed=4, spacing=3.0)
    out = {}
    for graph in (False, True):
        g = AgentGroup("twod", s0, P.InvPendulumBicycleParameters(), destqueues=list(queues_with_start(s0, q)),
                       dtype=torch.float32)
        eng = Engine([g], dtype=torch.float32, pair_mode="tiled", resort_every=16, graph=graph)
        for _ in range(40):
            eng.step()
        eng.check_status()
        out[graph] = (g.states_numpy(), eng.force.cpu().numpy())
    assert np.array_equal(out[False][0], out[True][0])
    assert np.array_equal(out[False][1], out[True][1])


def test_device_side_churn_select_concat():
    """AgentGroup.select / AgentGroup.concat (the device-side half of add_road_user / remove_road_user,
    reference intersection.py:458-539, :576-634): removing and re-adding far-away road users leaves the
    rest of a 2,500-cyclist crowd on the trajectory of a crowd that never changed, and the re-added users
    come back with their complete device record (navigation machine, history rings)."""
    from cyclistsocialforce_b200.engine import AgentGroup, Engine
    from cyclistsocialforce_b200.synthetic import queues_with_start
    n_main, n_far = 2500, 40
    s0, q = co.synthetic_crowd(n_main, seed=6, spacing=3.0)
    sf, qf = co.synthetic_crowd(n_far, seed=7, spacing=300.0)
    sf[:, 0] += 4000.0; qf[..., 0] += 4000.0
    S, Q = np.concatenate([s0, sf]), np.concatenate([q, qf])

    def group():
        return AgentGroup("twod", S, P.InvPendulumBicycleParameters(), destqueues=list(queues_with_start(S, Q)),
                          dtype=torch.float64)

    ga = group()
    ea = Engine([ga], dtype=torch.float64, pair_mode="tiled")
    for _ in range(30):
        ea.step()
    gb = group()
    eb = Engine([gb], dtype=torch.float64, pair_mode="tiled")
    for _ in range(10):
        eb.step()
    far = gb.select(np.arange(n_main, n_main + n_far))
    far_state = far.states_numpy()
    gm = gb.select(np.arange(n_main))
    assert gm.n == n_main and far.n == n_far
    eb = Engine([gm], dtype=torch.float64, pair_mode="tiled")
    for _ in range(10):
        eb.step()
    gc = AgentGroup.concat(gm, far)
    assert gc.n == n_main + n_far
    assert np.array_equal(gc.states_numpy()[n_main:], far_state)
    assert np.array_equal(gc.hist_step.cpu().numpy()[n_main:], far.hist_step.cpu().numpy())
    assert np.array_equal(gc.dest_ptr.cpu().numpy()[n_main:], far.dest_ptr.cpu().numpy())
    eb = Engine([gc], dtype=torch.float64, pair_mode="tiled")
    for _ in range(10):
        eb.step()
    eb.check_status()
    a, b = ga.states_numpy()[:n_main], gc.states_numpy()[:n_main]
    assert np.abs(a - b).max() < 1e-9
    assert np.array_equal(ga.dest_ptr.cpu().numpy()[:n_main], gc.dest_ptr.cpu().numpy()[:n_main])
