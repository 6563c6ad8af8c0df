"""ctypes binding of ``libcsf_b200.so`` (the C ABI declared in include/csf_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, an
exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# (CSF_B200_LIB: another build of the same library, for tuning experiments -- tools/gpu/*.sh)
LIB_PATH = os.environ.get("CSF_B200_LIB") or os.path.join(HERE, "libcsf_b200.so")

MODEL_IDS = dict(twod=0, invpendulum=1, balancingrider=2, planarpoint=3, bicycle=4)


class CsfFieldParams(C.Structure):
    _fields_ = [
        ("f_0", C.c_double), ("e_0", C.c_double), ("e_1", C.c_double),
        ("sigma_0", C.c_double), ("sigma_1", C.c_double), ("sigma_2", C.c_double), ("sigma_3", C.c_double),
        ("hfov", C.c_double), ("q_scale", C.c_double),
        ("p2r", C.c_int32), ("field_kind", C.c_int32),
        ("p_0", C.c_double), ("p_decay", C.c_double), ("v_max", C.c_double),
        ("cutoff_log2", C.c_double),
    ]


CSF_MAX_PEERS = 16
CSF_TILED_PREPARED, CSF_TILED_NO_REDUCE, CSF_TILED_PDL = 1, 2, 4


class CsfPeerComm(C.Structure):
    _fields_ = [
        ("world", C.c_int32), ("rank", C.c_int32),
        ("payload", C.c_void_p * CSF_MAX_PEERS),
        ("data_flags", C.c_void_p * CSF_MAX_PEERS),
        ("read_flags", C.c_void_p * CSF_MAX_PEERS),
        ("seq", C.c_void_p), ("status_host", C.c_void_p),
    ]


class CsfStepFusion(C.Structure):
    _fields_ = [
        ("partial", C.c_void_p), ("partial_stride", C.c_int64), ("partial_offset", C.c_int64),
        ("n_groups", C.c_int32), ("pdl", C.c_int32), ("f0", C.c_double), ("comm", CsfPeerComm),
    ]


class CsfAgentParams(C.Structure):
    _fields_ = [
        ("t_s", C.c_double),
        ("d_arrived_inter", C.c_double), ("d_arrived_stop", C.c_double),
        ("v_max_stop", C.c_double), ("v_max_harddecel", C.c_double),
        ("a_max", C.c_double * 2), ("a_desired", C.c_double * 2), ("v_max_riding", C.c_double * 2),
        ("l", C.c_double), ("delta_max", C.c_double), ("k_p_v", C.c_double), ("k_p_delta", C.c_double),
        ("g", C.c_double),
        ("l_2", C.c_double), ("tau_1_squared", C.c_double), ("i_steer", C.c_double), ("c_steer", C.c_double),
        ("v_max_walk", C.c_double), ("delta_max_walk", C.c_double),
        ("kx_table", (C.c_double * 4) * 5), ("ku_table", C.c_double * 4),
        ("k_psi", C.c_double),
        ("br_A0", C.c_double * 25), ("br_A1", C.c_double * 25), ("br_A2", C.c_double * 25),
        ("br_B", C.c_double * 5),
        ("br_pole_icpt", C.c_double * 5), ("br_pole_coef", C.c_double * 5),
        ("br_stochastic", C.c_int32), ("br_n_comp", C.c_int32), ("br_fixed_gains", C.c_int32), ("br_pad_", C.c_int32),
        ("br_resample_thresh", C.c_double),
        ("br_seed", C.c_uint64),
        ("br_lam", C.c_double * 6), ("br_sc_mean", C.c_double * 6), ("br_sc_scale", C.c_double * 6),
        ("br_log_a", C.c_double * 5), ("br_log_sign", C.c_double * 5),
        ("br_w", C.c_double * 4), ("br_mu_g", C.c_double * 4), ("br_var_g", C.c_double * 4),
        ("br_mu", (C.c_double * 5) * 4), ("br_slope", (C.c_double * 5) * 4), ("br_chol", (C.c_double * 15) * 4),
        ("q_scale", C.c_double), ("q_origin", C.c_double * 2),
        ("traj_len", C.c_int32), ("hist_len", C.c_int32), ("hist_cap", C.c_int32), ("q_cap", C.c_int32),
    ]


class CsfAgentState(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("first", C.c_int64), ("count", C.c_int64), ("payload_offset", C.c_int64),
        ("x", C.c_void_p), ("y", C.c_void_p),
        ("psi", C.c_void_p), ("v", C.c_void_p), ("delta", C.c_void_p), ("theta", C.c_void_p),
        ("deltadot", C.c_void_p), ("thetadot", C.c_void_p), ("vd_default", C.c_void_p),
        ("step_i", C.c_void_p),
        ("destq", C.c_void_p), ("dest_len", C.c_void_p), ("dest_ptr", C.c_void_p),
        ("znav", C.c_void_p), ("znav_v0", C.c_void_p), ("znav_d0", C.c_void_p), ("znav_d1", C.c_void_p),
        ("prev_x", C.c_void_p), ("prev_y", C.c_void_p), ("hist_x", C.c_void_p), ("hist_y", C.c_void_p),
        ("hist_step", C.c_void_p),
        ("ip_x", C.c_void_p), ("ip_zrid", C.c_void_p), ("ip_delta_run", C.c_void_p),
        ("dyn_x", C.c_void_p), ("dyn_v", C.c_void_p), ("br_gains", C.c_void_p),
        ("br_poles", C.c_void_p), ("br_vlast", C.c_void_p), ("br_draws", C.c_void_p), ("br_stream", C.c_void_p),
        ("status", C.c_void_p), ("status_host", C.c_void_p),
    ]


_vp, _i64, _i32, _dbl, _sz = C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_size_t
_FP = C.POINTER(CsfFieldParams)
MAX_COPY_SEGMENTS = 16


class CsfCopySegment(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("bytes", C.c_int64)]


class CsfCopySegments(C.Structure):
    _fields_ = [("n", C.c_int32), ("pad_", C.c_int32), ("seg", CsfCopySegment * MAX_COPY_SEGMENTS)]


MAX_GATHER_SEGMENTS = 48


class CsfGatherSegment(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("idx", C.c_void_p), ("outer", C.c_int64), ("n_src", C.c_int64),
                ("n_dst", C.c_int64), ("dst_off", C.c_int64), ("count", C.c_int64), ("inner_bytes", C.c_int64),
                ("dst_inner_bytes", C.c_int64)]


class CsfGatherSegments(C.Structure):
    _fields_ = [("n", C.c_int32), ("pad_", C.c_int32), ("seg", CsfGatherSegment * MAX_GATHER_SEGMENTS)]


_AP = C.POINTER(CsfAgentParams)
_AS = C.POINTER(CsfAgentState)

#: every symbol include/csf_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "csf_version": (C.c_int, []),
    "csf_last_error_string": (C.c_char_p, []),
    "csf_sm_count": (C.c_int, []),
    "csf_pair_workspace_bytes": (_sz, [_i64, _i64, C.c_int]),
    "csf_pair_forces_f32": (C.c_int, [_vp, _i64, _vp, _i64, _FP, _vp, C.c_int, _vp, _sz, _vp]),
    "csf_pair_forces_f64": (C.c_int, [_vp, _i64, _vp, _i64, _FP, _vp, C.c_int, _vp, _sz, _vp]),
    "csf_pair_forces_grouped_f32": (C.c_int, [_vp, _i64, _i32, _FP, _vp, _vp]),
    "csf_pair_forces_grouped_f64": (C.c_int, [_vp, _i64, _i32, _FP, _vp, _vp]),
    "csf_pair_forces_bicycle_f32": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _FP, _vp, C.c_int, _vp]),
    "csf_pair_forces_bicycle_f64": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _FP, _vp, C.c_int, _vp]),
    "csf_bicycle_eccentricity_f32": (C.c_int, [_vp, _i64, _dbl, _vp, _vp]),
    "csf_bicycle_eccentricity_f64": (C.c_int, [_vp, _i64, _dbl, _vp, _vp]),
    "csf_tiled_padded_sources": (_i64, [_i64]),
    "csf_tiled_num_tiles": (_i64, [_i64]),
    "csf_tiled_tile_bytes": (C.c_int, [C.c_int]),
    "csf_pair_tiled_workspace_bytes": (_sz, [_i64, _i64, C.c_int]),
    "csf_morton_keys_f32": (C.c_int, [_vp, _i64, _dbl, _dbl, _dbl, _vp, _vp]),
    "csf_morton_keys_f64": (C.c_int, [_vp, _i64, _dbl, _dbl, _dbl, _vp, _vp]),
    "csf_spatial_keys_f32": (C.c_int, [_vp, _i64, _vp, _vp, _vp]),
    "csf_spatial_keys_f64": (C.c_int, [_vp, _i64, _vp, _vp, _vp]),
    "csf_spatial_order_workspace_bytes": (_sz, [_i64]),
    "csf_spatial_bbox_f32": (C.c_int, [_vp, _i64, _vp, _vp]),
    "csf_spatial_bbox_f64": (C.c_int, [_vp, _i64, _vp, _vp]),
    "csf_spatial_order_f32": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "csf_spatial_order_f64": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "csf_tile_sources_f32": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "csf_tile_sources_f64": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "csf_tile_sources_bicycle_f32": (C.c_int, [_vp, _vp, _dbl, _i64, _vp, _vp, _vp, _vp]),
    "csf_tile_sources_bicycle_f64": (C.c_int, [_vp, _vp, _dbl, _i64, _vp, _vp, _vp, _vp]),
    "csf_pair_forces_tiled_f32": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _FP, _vp, C.c_int, _vp, _sz, _vp, _vp, _vp, C.c_int, _vp]),
    "csf_pair_forces_tiled_f64": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _FP, _vp, C.c_int, _vp, _sz, _vp, _vp, _vp, C.c_int, _vp]),
    "csf_tiled_prepare_f32": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _sz, _vp, _vp]),
    "csf_tiled_prepare_f64": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _sz, _vp, _vp]),
    "csf_tiled_num_groups": (C.c_int, [_i64, _i64, C.c_int]),
    "csf_tiled_partial_offset": (_sz, [_i64, _i64, C.c_int]),
    "csf_tiled_num_items": (_i64, [_i64, _i64, C.c_int]),
    "csf_tiled_item_order": (C.c_int, [_vp, _i64, _vp, _vp]),
    "csf_field_cutoff_distance": (_dbl, [_FP]),
    "csf_field_reach_table": (C.c_int, [_FP, C.c_int, C.POINTER(C.c_double)]),
    "csf_road_forces_f32": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _dbl, _dbl, _vp, C.c_int, _vp]),
    "csf_road_forces_f64": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _dbl, _dbl, _vp, C.c_int, _vp]),
    "csf_agent_forces_f32": (C.c_int, [C.c_int, _AS, _AP, _i64, _vp, _vp, _vp, _vp, _vp]),
    "csf_agent_forces_f64": (C.c_int, [C.c_int, _AS, _AP, _i64, _vp, _vp, _vp, _vp, _vp]),
    "csf_agent_advance_f32": (C.c_int, [C.c_int, _AS, _AP, _vp, _vp, _vp]),
    "csf_agent_advance_f64": (C.c_int, [C.c_int, _AS, _AP, _vp, _vp, _vp]),
    "csf_agent_step_f32": (C.c_int, [C.c_int, _AS, _AP, _i64, _vp, _vp, _vp, _vp, _vp]),
    "csf_agent_step_f64": (C.c_int, [C.c_int, _AS, _AP, _i64, _vp, _vp, _vp, _vp, _vp]),
    "csf_agent_step_fused_f32": (C.c_int, [C.c_int, _AS, _AP, _i64, _vp, _vp, _vp, _vp, _vp]),
    "csf_agent_step_fused_f64": (C.c_int, [C.c_int, _AS, _AP, _i64, _vp, _vp, _vp, _vp, _vp]),
    "csf_br_init": (C.c_int, [_AS, _AP, _vp]),
    "csf_pack_xycs_f32": (C.c_int, [_AS, _AP, _vp, _vp]),
    "csf_pack_xycs_f64": (C.c_int, [_AS, _AP, _vp, _vp]),
    "csf_pack_xypsi_f32": (C.c_int, [_vp, _vp, _vp, _i64, _dbl, _dbl, _dbl, _vp, _vp]),
    "csf_pack_xypsi_f64": (C.c_int, [_vp, _vp, _vp, _i64, _dbl, _dbl, _dbl, _vp, _vp]),
    "csf_copy_segments": (C.c_int, [C.POINTER(CsfCopySegments), _vp]),
    "csf_gather_segments": (C.c_int, [C.POINTER(CsfGatherSegments), _vp]),
    "csf_sumo_pose_f32": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "csf_sumo_pose_f64": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "csf_ffma_peak": (C.c_int, [_i64, _vp, C.POINTER(C.c_double), _vp]),
    "csf_peer_handle_bytes": (C.c_int, []),
    "csf_peer_alloc": (C.c_int, [_sz, C.POINTER(C.c_void_p), _vp]),
    "csf_peer_open": (C.c_int, [_vp, C.POINTER(C.c_void_p)]),
    "csf_peer_close": (C.c_int, [_vp]),
    "csf_peer_free": (C.c_int, [_vp]),
    "csf_peer_wait_data": (C.c_int, [C.POINTER(CsfPeerComm), _vp]),
    "csf_peer_signal_read": (C.c_int, [C.POINTER(CsfPeerComm), _vp]),
    "csf_peer_push": (C.c_int, [C.POINTER(CsfPeerComm), _i64, _i64, C.c_int, _vp]),
}

_lib = None


class CsfError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it is missing -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CsfError(
            f"{LIB_PATH} not found: build it with `python -m cyclistsocialforce_b200.build` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.csf_version() != 1:
        raise CsfError("libcsf_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().csf_last_error_string().decode(errors="replace")
        raise CsfError(f"{what} failed (rc={rc}): {msg}")
