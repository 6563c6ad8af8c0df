"""Generate tests/golden/golden_v1.npz from the REFERENCE's own code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the unmodified reference through ``oracle/ref_harness.py`` (stubs for
presentation packages, restated substitutes for python-control and
bicycleparameters, constructor fix for TwoDBicycle -- see that file) and records
inputs and outputs of ``SocialForceIntersection.calc_forces()/step()``
(reference intersection.py:747-896) on small seeded cases.  The vectors are the
pin for ``oracle/csf_oracle.py`` and, through it, for the CUDA path.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from oracle.csf_oracle import synthetic_crowd  # noqa: E402

veh, ins_m, par, dyn, utils = rh.modules()

DEMO_S0 = np.array([(-6, 0, 0, 5, 0, 0, 0, 0), (15, -20, np.pi / 2, 5, 0, 0, 0, 0),
                    (13, -20, np.pi / 2, 5, 0, 0, 0, 0)], float)
DEMO_VD = np.array([4.5, 5.0, 5.0])
DEMO_DX = np.array([(35, 64, 65), (15, 15, 15), (13, 13, 13)], float)
DEMO_DY = np.array([(0, 0, 0), (20, 49, 50), (20, 49, 50)], float)
CLASSES = dict(twod=veh.TwoDBicycle, invpendulum=veh.InvPendulumBicycle,
               balancingrider=veh.BalancingRiderBicycle, planarpoint=veh.PlanarPointBicycle,
               bicycle=veh.Bicycle)
out = {}


def build(cls, s0, vd, dests, **ins_kw):
    bikes = []
    for k in range(s0.shape[0]):
        b = cls(tuple(s0[k, :cls.N_STATES]), id=str(k))
        b.params.v_desired_default = float(vd[k])
        d = np.asarray(dests[k], float)
        b.setDestinations(d[:, 0], d[:, 1], stop=d[:, 2] if d.shape[1] > 2 else None)
        bikes.append(b)
    return rh.headless_intersection(bikes, **ins_kw)


def record(tag, ins, steps, keep):
    S, F = [], []
    for k in range(1, steps + 1):
        ins.step()
        if k in keep:
            S.append(np.array([v.s for v in ins.vehicles]))
            F.append(np.array([v.force for v in ins.vehicles]))
    out[tag + "_steps"] = np.array(sorted(keep))
    out[tag + "_s"] = np.array(S)
    out[tag + "_F"] = np.array(F)
    out[tag + "_ptr"] = np.array([v.destpointer for v in ins.vehicles])
    out[tag + "_znav"] = np.array([v.znav for v in ins.vehicles])


# 1. demo geometry (demo/demoCSFstandalone.py:101-118), one run per model class
demo_dests = np.stack([DEMO_DX, DEMO_DY, np.zeros_like(DEMO_DX)], axis=2)
out["demo_s0"], out["demo_vd"], out["demo_dests"] = DEMO_S0, DEMO_VD, demo_dests
for name, cls in CLASSES.items():
    steps, keep = (300, {1, 10, 100, 300}) if name == "balancingrider" else \
                  (1500 if name == "twod" else 700, {1, 10, 100, 400, 700})
    if name == "twod":
        keep |= {1100, 1500}
    record("demo_" + name, build(cls, DEMO_S0, DEMO_VD, demo_dests), steps, keep)

# 2. stop destinations (nav state machine, vehicle.py:354-457): last dest is a stop
stop_dests = demo_dests.copy()
stop_dests[:, 2, 2] = 1.0
out["stop_dests"] = stop_dests
for name in ("twod", "invpendulum", "bicycle"):   # planarpoint: reference raises TypeError at vehicle.py:1556
    record("stop_" + name, build(CLASSES[name], DEMO_S0, DEMO_VD, stop_dests), 2200,
           {100, 900, 1400, 1500, 1600, 1700, 1800, 2000, 2200})

# 3. priority to the right (intersection.py:738-741)
record("p2r_twod", build(veh.TwoDBicycle, DEMO_S0, DEMO_VD, demo_dests, priority_rule="p2r"),
       700, {1, 100, 400, 700})

# 4. seeded synthetic crowd: N=24, spacing 3, seed 3
s0, q = synthetic_crowd(24, seed=3, spacing=3.0)
out["crowd_s0"], out["crowd_dests"] = s0, q
ins = build(veh.TwoDBicycle, s0, np.full(24, 5.0), q)
record("crowd_twod", ins, 60, {1, 5, 30, 60})
# full pair matrix and mask at the final state
tracked = ~ins.get_untracked_foes()
n = ins.n_bikes
Fx = np.zeros((n, n)); Fy = np.zeros((n, n))
for i in range(n):
    m = tracked[i]
    fx, fy = ins.vehicles[i].calcRepulsiveForce(ins.vehicleX[m, 0], ins.vehicleY[m, 0],
                                                ins.vehicleTheta[m, 0])
    Fx[i, m], Fy[i, m] = fx, fy
out["crowd_pair_xypsi"] = np.c_[ins.vehicleX, ins.vehicleY, ins.vehicleTheta]
out["crowd_pair_tracked"], out["crowd_pair_Fx"], out["crowd_pair_Fy"] = tracked, Fx, Fy

# 5. road-edge force (intersection.py:226-242) on the curve-scenario geometry
seg1 = ins_m.StraightRoadSegment(np.array([0.0, 0.0, np.pi / 2]), 3.0, 20.0,
                                 params=par.RoadElementParameters(F_0=0.15, sigma=2.0))
seg2 = ins_m.CurvedRoadSegment(seg1.x1, 3.0, 8.0, np.pi / 2, "right",
                               params=par.RoadElementParameters(F_0=0.15, sigma=2.0))
col = ins_m.RoadSegmentCollection([seg1, seg2])
rng = np.random.default_rng(7)
px = rng.uniform(-1.2, 1.2, 16); py = rng.uniform(0, 20, 16)
fx, fy = col.calcRepulsiveForce(px, py)
out["road_pts"] = np.c_[px, py]
out["road_F"] = np.c_[fx, fy]
for k, e in enumerate(seg1.edges + seg2.edges):
    out[f"road_edge{k}"] = e.vertices
out["road_x1"] = np.array([seg1.x1, seg2.x1])

# 6. utils (utils.py:124-194)
ang = np.r_[np.linspace(-10, 10, 41), np.pi, -np.pi, 3 * np.pi, 0.0]
out["util_angles"] = ang
out["util_limit"] = np.array([utils.limitAngle(float(a)) for a in ang])
A1, A2 = np.meshgrid(ang[::3], ang[1::3])
out["util_a1"], out["util_a2"] = A1.ravel(), A2.ravel()
out["util_angdiff"] = np.array([utils.angleDifference(float(a), float(b))
                                for a, b in zip(A1.ravel(), A2.ravel())])

# 7. InvPendulum closed loop + gains; BalancingRider matrices/gains at v = 5
p = par.InvPendulumBicycleParameters()
kx, ku = p.fullstate_feedback_gains(5.0)
out["invpend_kx_v5"], out["invpend_ku_v5"] = kx.ravel(), np.array(ku)
b = veh.BalancingRiderBicycle((0, 0, 0, 5, 0, 0, 0, 0))
A, B, _, _ = b.dynamics.get_statespace_matrices(5.0)
out["br_A_v5"], out["br_B_v5"] = A, B[:, 1]
out["br_gains_v5"] = np.asarray(b.dynamics._get_gains(5.0)).ravel()
out["br_poles_v5"] = np.array(b.params.poles)
M, C1, K0, K2 = b.params.bp_model.form_reduced_canonical_matrices()
out["br_M"], out["br_C1"], out["br_K0"], out["br_K2"] = M, C1, K0, K2

# 8. parcours scenario (scenarios/parcours-scenario.py:31-40), single BalancingRider
bb = veh.BalancingRiderBicycle((0, 0, np.pi / 2, 5, 0, 0, 0, 0), id="p")
bb.params.v_desired_default = 4.0
pdx = [0, 10, 0, 5, 10, 20, 21, 22, 23]; pdy = [10, 20, 30, 40, 40, 40, 40, 40, 40]
bb.setDestinations(pdx, pdy)
pins = rh.headless_intersection([])
pins.add_road_user(bb)
out["parcours_dests"] = np.c_[pdx, pdy]
record("parcours", pins, 1500, {1, 100, 500, 1000, 1500})

path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")
