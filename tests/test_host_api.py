"""CPU-only tests: the C-ABI library loads and exports every declared symbol, the host
mirror of the reference interface behaves like the reference, and the product path fails
loudly without a GPU (no fallback)."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import csf_oracle as co
from cyclistsocialforce_b200 import _lib, parameters as P, vehicle as V
from cyclistsocialforce_b200 import intersection as I
from cyclistsocialforce_b200.scenario import Scenario

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAS_CUDA = torch.cuda.is_available()


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "csf_b200.h")).read()
    declared = set(re.findall(r"\b(csf_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.csf_version() == 1


def test_struct_layout_matches_header():
    """ctypes mirrors vs the C compiler's sizeof (gcc on the header)."""
    import subprocess, tempfile, ctypes
    src = ('#include "csf_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n",'
           'sizeof(CsfAgentParams),sizeof(CsfAgentState),sizeof(CsfFieldParams));return 0;}')
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", os.path.join(d, "s")])
        out = subprocess.check_output([os.path.join(d, "s")]).decode().split()
    assert [int(x) for x in out] == [ctypes.sizeof(_lib.CsfAgentParams), ctypes.sizeof(_lib.CsfAgentState),
                                     ctypes.sizeof(_lib.CsfFieldParams)]


@pytest.mark.parametrize("model,cls", [("twod", P.InvPendulumBicycleParameters),
                                       ("balancingrider", P.BalancingRiderBicycleParameters),
                                       ("planarpoint", P.PlanarPointBicycleParameters),
                                       ("bicycle", P.BicycleParameters), ("uncontrolled", P.CarParameters)])
def test_parameter_defaults_match_reference(model, cls):
    """Defaults equal the oracle's table, which is pinned to the reference (Appendix C)."""
    ref = co.default_params(model)
    p = cls()
    for name, val in vars(ref).items():
        if name in ("bike", "pole_model", "k_psi", "tau_1_squared"):
            continue
        got = getattr(p, name)
        assert np.allclose(np.asarray(got, float), np.asarray(val, float), rtol=0, atol=0), name
    if model == "twod":
        assert p.tau_1_squared == ref.tau_1_squared
        ap = p.to_agent_params(2.0 ** -20, 6, 128)
        assert ap.traj_len == 3000 and ap.hist_len == 100
        kx, ku = co.invpend_gains(4.0)
        assert np.allclose(p.fullstate_feedback_gains(4.0)[0].ravel(), kx) and p.fullstate_feedback_gains(4.0)[1] == ku


def test_balancingrider_constants_match_oracle():
    p = P.BalancingRiderBicycleParameters()
    ap = p.to_agent_params(1.0, 4, 128)
    for v in (1.5, 4.0, 6.5):
        A, B = co.balancingrider_matrices(co.BALANCEASSIST, v)
        Ak = (np.array(ap.br_A0) + v * np.array(ap.br_A1) + v * v * np.array(ap.br_A2)).reshape(5, 5)
        assert np.abs(Ak - A).max() < 1e-12
        assert np.abs(np.array(ap.br_B) - B).max() < 1e-15
        assert np.allclose(np.sort_complex(np.array(p.poles_at(v))),
                           np.sort_complex(co.balancingrider_poles(("BR1", 0), v)))
    assert abs(p.l - 1.113) < 1e-15 and abs(p.m - 102.12) < 1e-9


def test_parameter_error_behaviour():
    with pytest.raises(TypeError):
        P.RoadElementParameters(F_0=1)           # reference: "F_0 must be a float."
    with pytest.raises(ValueError):
        P.RoadElementParameters(sigma=-1.0)
    r = P.RoadElementParameters()
    with pytest.raises(AttributeError):
        r.F_0 = 0.3                               # immutable (parameters.py:385-386)
    with pytest.raises(TypeError):
        V.TwoDBicycle((0, 0, 0, 5, 0), params=P.BicycleParameters())
    with pytest.raises(ValueError):
        V.TwoDBicycle((0, 0, 0))                  # too few states (vehicle.py:149-150)
    with pytest.raises(NotImplementedError):
        V.Vehicle((0, 0, 0, 1), rep_force_func=lambda *a: (0, 0))


def test_vehicle_host_logic():
    b = V.TwoDBicycle((1, 2, 7.0, 5, 0.1, 9, 9), id="a")       # extra states are cut (vehicle.py:151-152)
    assert b.s.shape == (5,) and abs(b.s[2] - co.limit_angle(7.0)) < 1e-15
    assert b.N_STATES == 5 and b.STATE_NAMES[4] == "delta[rad]"
    assert np.array_equal(b.destqueue, [[1, 2, 0]])            # vehicle.py:183-185
    b.setDestinations((10, 20), (0, 5))
    assert b.destqueue.shape == (3, 3) and not b.isLastDest()
    b.setDestinations(3.0, 4.0, stop=1.0, reset=True)
    assert np.array_equal(b.destqueue, [[3, 4, 1]]) and b.isLastDest() and b.destpointer == 0
    assert abs(b.getDestinationDistance() - np.hypot(2, 2)) < 1e-15
    assert b.traj.shape == (5, 3000) and np.array_equal(b.traj[:, 0], b.s)
    assert V.InvertedPendulumBicycle is V.InvPendulumBicycle
    ip = V.InvPendulumBicycle((0, 0, 0, 1.0, 0, 0))
    assert ip.zrid.tolist() == [False, True]                   # starts walking below v_max_walk
    u = V.UncontrolledVehicle((0, 0, 0, 0), trajectory=np.array([[0, 1, 2], [0, 0, 0], [0, 0, 0], [1, 1, 1]]))
    u.step()
    assert u.s[0] == 1.0
    assert u.calcDestinationForce() == (0, 0)


def test_road_geometry_matches_reference(golden):
    seg1 = I.StraightRoadSegment(np.array([0.0, 0.0, np.pi / 2]), 3.0, 20.0,
                                 params=P.RoadElementParameters(F_0=0.15, sigma=2.0))
    seg2 = I.CurvedRoadSegment(seg1.x1, 3.0, 8.0, np.pi / 2, "right",
                               params=P.RoadElementParameters(F_0=0.15, sigma=2.0))
    col = I.RoadSegmentCollection([seg1, seg2])
    flat = col.edges_flat()
    assert len(flat) == 4
    for k, (verts, F_0, sigma) in enumerate(flat):
        assert np.abs(verts - golden[f"road_edge{k}"]).max() < 1e-13
        assert (F_0, sigma) == (0.15, 2.0)
    assert np.abs(np.array([seg1.x1, seg2.x1]) - golden["road_x1"]).max() < 1e-13


def test_scenario_loop():
    calls = []
    scn = Scenario(lambda: calls.append(1), t_s=0.01, t_r=0.0, verbose=False)
    scn.run(0.25)
    assert len(calls) == 25 and scn.i == 25 and abs(scn.t - 0.25) < 1e-12
    scn.reset()
    assert scn.i == 0 and scn.t == 0
    with pytest.raises(NotImplementedError):
        Scenario(lambda: None, animate=True)


@pytest.mark.skipif(HAS_CUDA, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    b = V.TwoDBicycle((0, 0, 0, 5, 0))
    with pytest.raises(_lib.CsfError):
        I.SocialForceIntersection([b])
    with pytest.raises(_lib.CsfError):
        b.calcDestinationForce()


def test_q_scale_choice():
    q = P.choose_q_scale(5000.0)
    assert 5000.0 / q <= 2 ** 30 < 2 * 5000.0 / q * 1.0000001 * 2
    assert np.log2(q) == np.floor(np.log2(q))


def test_payload_frame_is_centred_on_the_crowd():
    """The Q-format frame of the f32 payload: origin = centre of the bounding box of everything the crowd
    can reach, extent = half its size plus a margin -- the 31 bits cover the occupied region only."""
    rng = np.random.default_rng(0)
    pos = rng.uniform(1000.0, 1256.0, (500, 2))
    dest = rng.uniform(900.0, 1400.0, (500, 5, 2))
    (ox, oy), ext = P.payload_frame([pos, dest])
    lo = np.minimum(pos.min(0), dest.reshape(-1, 2).min(0))
    hi = np.maximum(pos.max(0), dest.reshape(-1, 2).max(0))
    assert abs(ox - (lo[0] + hi[0]) / 2) < 1e-9 and abs(oy - (lo[1] + hi[1]) / 2) < 1e-9
    assert np.all(np.abs(pos - (ox, oy)) < ext) and np.all(np.abs(dest - (ox, oy)) < ext)
    assert ext < 0.2 * (2 * 1400.0 + 1000.0)           # far tighter than a frame around (0, 0)
    assert P.choose_q_scale(ext) <= 2.0 ** -20
    assert P.payload_frame([]) == ((0.0, 0.0), 1000.0)


def test_package_utils_match_the_reference_vectors(golden):
    """cyclistsocialforce_b200.utils (host helpers with the reference's names) against vectors generated by
    the reference's own utils (tests/golden/make_golden.py): bit for bit, scalar and array call forms."""
    from cyclistsocialforce_b200 import utils as U
    g = golden
    assert np.array_equal(U.limitAngle(g["util_angles"].copy()), g["util_limit"])
    assert [U.limitAngle(float(a)) for a in g["util_angles"]] == g["util_limit"].tolist()
    assert isinstance(U.limitAngle(0.3), float)
    assert np.array_equal(U.angleDifference(g["util_a1"].copy(), g["util_a2"].copy()), g["util_angdiff"])
    assert [U.angleDifference(float(a), float(b)) for a, b in zip(g["util_a1"], g["util_a2"])] == g["util_angdiff"].tolist()
    rho, phi = U.cart2polar(np.array([1.0, 0.0, -2.0, 3.0]), np.array([0.0, 2.0, -0.0, -4.0]))
    assert np.allclose(rho, [1, 2, 2, 5]) and np.allclose(phi, [0, np.pi / 2, np.pi, -np.arccos(0.6)])
    assert np.array_equal(U.thresh(np.array([-2.0, 0.5, 9.0]), (-1.0, 1.0)), [-1.0, 0.5, 1.0])
    with pytest.raises(AssertionError):
        U.thresh(1.0, (2.0, 1.0))
    x, y = np.array([3.0, 0.3, 0.0]), np.array([4.0, 0.4, 0.0])
    rx, ry = U.limitMagnitude(x, y, np.array([1.0, 1.0, 1.0]))
    assert rx is x and ry is y                                  # in place, like the reference
    assert np.allclose(x, [0.6, 0.3, 0.0]) and np.allclose(y, [0.8, 0.4, 0.0])
    z = np.zeros(3)
    U.limitMagnitude(z, z.copy(), np.ones(3))
    assert not z.any()


@pytest.mark.reference
def test_package_utils_match_the_reference_live():
    from oracle import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("no /root/reference")
    from cyclistsocialforce_b200 import utils as U
    ref = rh.modules()[4]
    rng = np.random.default_rng(3)
    a1, a2 = rng.uniform(-4 * np.pi, 4 * np.pi, 2000), rng.uniform(-np.pi, np.pi, 2000)
    assert np.array_equal(U.limitAngle(a1.copy()), ref.limitAngle(a1.copy()))
    w1 = ref.limitAngle(a1.copy())
    assert np.array_equal(U.angleDifference(w1.copy(), a2.copy()), ref.angleDifference(w1.copy(), a2.copy()))
    for p, q in zip(w1[:200], a2[:200]):
        assert U.angleDifference(float(p), float(q)) == ref.angleDifference(float(p), float(q))
    x, y = rng.normal(size=500), rng.normal(size=500)
    for got, want in zip(U.cart2polar(x, y), ref.cart2polar(x, y)):
        assert np.array_equal(got, want)
    r = np.abs(rng.normal(size=500))
    gx, gy = U.limitMagnitude(x.copy(), y.copy(), r)
    wx, wy = ref.limitMagnitude(x.copy(), y.copy(), r)
    assert np.array_equal(gx, wx) and np.array_equal(gy, wy)
    assert np.array_equal(U.thresh(x, (-0.5, 0.7)), ref.thresh(x, (-0.5, 0.7)))


@pytest.mark.reference
@pytest.mark.parametrize("direction", ["left", "right"])
def test_road_geometry_matches_reference_live(direction):
    """Vertices of straight and curved segments (both turn directions, a rotated start pose) against the
    reference's own classes."""
    from oracle import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("no /root/reference")
    R = rh.modules()[1]
    x0 = np.array([3.0, -2.0, 0.7])
    mine1, ref1 = I.StraightRoadSegment(x0, 2.5, 12.3), R.StraightRoadSegment(x0, 2.5, 12.3)
    mine2 = I.CurvedRoadSegment(mine1.x1, 2.5, 6.0, 1.1, direction)
    ref2 = R.CurvedRoadSegment(ref1.x1, 2.5, 6.0, 1.1, direction)
    for m, r in ((mine1, ref1), (mine2, ref2)):
        assert np.abs(np.asarray(m.x1, float) - np.asarray(r.x1, float)).max() < 1e-12
        for em, er in zip(m.edges, r.edges):
            assert em.vertices.shape == np.asarray(er.vertices).shape
            assert np.abs(em.vertices - np.asarray(er.vertices)).max() < 1e-12


def test_empty_intersection_steps_without_a_device():
    """An intersection without road users (reference :866-896 with n_bikes == 0: nothing to do but the
    ``hist_n_vecs`` bookkeeping) needs no device and returns empty force arrays."""
    from cyclistsocialforce_b200.intersection import SocialForceIntersection
    ins = SocialForceIntersection([])
    ins.step()
    ins.step()
    fx, fy = ins.calc_forces()
    assert fx.shape == (0,) and fy.shape == (0,)
    assert ins.n_bikes == 0 and ins.hist_n_vecs == [0, 0] and ins.get_road_user_ids() == []
    assert ins.check_status() == []


def test_sumo_cosimulation_needs_a_client():
    """activate_sumo_cosimulation=True without traci and without a client object: a clear ImportError (the
    reference imports traci at module level, scenario.py:33-50)."""
    import sys
    from cyclistsocialforce_b200.intersection import SocialForceIntersection
    try:
        import traci  # noqa: F401  (absent from this image; other tests of the session may have stubbed it)
        have = True
    except ImportError:
        have = False
    if not have and "traci" not in sys.modules:
        with pytest.raises(ImportError):
            SocialForceIntersection([], activate_sumo_cosimulation=True)
    ins = SocialForceIntersection([], activate_sumo_cosimulation=True, sumo_client=object())
    assert ins.activate_sumo_cosimulation
    ins.step()                                   # no road users: nothing is handed over


def test_copy_segment_struct_matches_header():
    """CsfCopySegments (trajectory stream) as ctypes == the header's layout: 16 segments of {src, dst, bytes}."""
    import ctypes as C
    from cyclistsocialforce_b200 import _lib
    assert C.sizeof(_lib.CsfCopySegment) == 24
    assert C.sizeof(_lib.CsfCopySegments) == 8 + 16 * 24
    assert _lib.MAX_COPY_SEGMENTS == 16
    hdr = open(os.path.join(ROOT, "include", "csf_b200.h")).read()
    assert "#define CSF_MAX_COPY_SEGMENTS 16" in hdr
    # churn: CsfGatherSegments (three pointers + seven 64-bit integers per segment), passed by value to the
    # kernel: must stay below the 4 KB that every CUDA 12 driver accepts for kernel parameters
    assert C.sizeof(_lib.CsfGatherSegment) == 80
    assert C.sizeof(_lib.CsfGatherSegments) == 8 + _lib.MAX_GATHER_SEGMENTS * 80 <= 4096
    assert f"#define CSF_MAX_GATHER_SEGMENTS {_lib.MAX_GATHER_SEGMENTS}" in hdr


def test_trajectory_chunk_decoding():
    """Host side of the trajectory stream: a drained chunk (bytes of [state slabs | forces] per step) decodes
    into per-column arrays -- mixed double / float columns at 256-byte aligned offsets, two groups."""
    from cyclistsocialforce_b200.trajstream import decode_chunk
    rng = np.random.default_rng(0)
    steps, na, nb = 5, 7, 3
    lay_a = {"x": (0, na, np.float64), "y": (256, na, np.float64), "psi": (512, na, np.float32), "v": (768, na, np.float32)}
    lay_b = {"x": (0, nb, np.float64), "y": (256, nb, np.float64), "psi": (512, nb, np.float32)}
    off_a, off_b, force_off = 0, 1024, 2048
    force_bytes = (na + nb) * 2 * 4
    slot = force_off + 256
    buf = np.zeros((steps, slot), dtype=np.uint8)
    truth = {}
    for gi, (off, lay) in enumerate(((off_a, lay_a), (off_b, lay_b))):
        for name, (o, n, dt) in lay.items():
            val = rng.normal(size=(steps, n)).astype(dt)
            truth[(gi, name)] = val
            buf[:, off + o: off + o + n * np.dtype(dt).itemsize] = val.view(np.uint8).reshape(steps, -1)
    f = rng.normal(size=(steps, na + nb, 2)).astype(np.float32)
    buf[:, force_off: force_off + force_bytes] = f.reshape(steps, -1).view(np.uint8)
    rec = decode_chunk(buf, [(off_a, lay_a), (off_b, lay_b)], force_off, force_bytes, np.float32)
    assert rec["steps"] == steps and np.array_equal(rec["force"], f)
    for (gi, name), val in truth.items():
        assert np.array_equal(rec["groups"][gi][name], val)
