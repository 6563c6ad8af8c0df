/*
 * csf_b200.h -- C ABI of the B200-native social-force stepping engine.
 *
 * The reference (chris-konrad/cyclistsocialforce, pure Python) has no FFI layer;
 * this header is the boundary a replacement shared library exports for the one
 * hot path  SocialForceIntersection.step()  (reference
 * src/cyclistsocialforce/intersection.py:866-896).  Each entry point names the
 * reference code it replaces.  All functions are asynchronous on `stream`,
 * allocate nothing, take plain device pointers owned by the caller, and return
 * 0 on success or a negative cudaError_t.  One scalar type per entry point:
 *   _f32  production build   (pair payload: Q-format int32 positions + fp32 cos/sin)
 *   _f64  verification build (everything double)
 *
 * Layouts
 * -------
 *  pair payload ("xycs"), one element per road user, array [N]:
 *     f32: { int32 xq, yq; float cos_psi, sin_psi; }   16 B,  x = q_origin[0] + xq * q_scale  [m]
 *     f64: { double x, y, cos_psi, sin_psi; }          32 B
 *  forces: [n][2] (x,y interleaved) of the scalar type.
 *  per-agent state: struct of arrays, see CsfAgentState.
 */
#ifndef CSF_B200_H
#define CSF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSF_ABI_VERSION 1
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef void* csf_stream_t; /* cudaStream_t */

/* model ids (reference classes in src/cyclistsocialforce/vehicle.py) */
enum {
    CSF_MODEL_TWOD = 0,           /* TwoDBicycle            :1292-1648 */
    CSF_MODEL_INVPENDULUM = 1,    /* InvPendulumBicycle     :1651-1950 */
    CSF_MODEL_BALANCINGRIDER = 2, /* BalancingRiderBicycle  :1953-1988 + dynamics.py:261-705 */
    CSF_MODEL_PLANARPOINT = 3,    /* PlanarPointBicycle     :1991-2028 + dynamics.py:802-1079 */
    CSF_MODEL_BICYCLE = 4,        /* Bicycle (v0.1)         :990-1289 */
    CSF_MODEL_COUNT = 5
};

/* Repulsive force-field parameters of one class of *source* road users
 * (VehicleParameters f_0,e_0,e_1,sigma_0..3,hfov: parameters.py:430-451;
 * the reference reads them from the source vehicle, vehicle.py:1587-1612,
 * intersection.py:733-735). */
typedef struct CsfFieldParams {
    double f_0, e_0, e_1;
    double sigma_0, sigma_1, sigma_2, sigma_3;
    double hfov;     /* full horizontal field of view [rad] */
    double q_scale;  /* metres per position unit of the f32 payload (ignored by _f64) */
    int32_t p2r;     /* priority_rule == "p2r" (intersection.py:738-741) */
    int32_t field_kind; /* 0: TwoDBicycle field (vehicle.py:1560-1648); 1: Bicycle v0.1 (:1107-1147) */
    double p_0, p_decay, v_max; /* field_kind 1 only (BicycleParameters p_0, p_decay, v_max_riding[1]) */
    double cutoff_log2; /* tiled f32 kernel: contributions below 2^-cutoff_log2 * f_0 may be dropped (0 = default 40) */
} CsfFieldParams;

/* Crowd-wide parameters of the per-agent kernels (all models share one struct). */
typedef struct CsfAgentParams {
    double t_s;
    double d_arrived_inter, d_arrived_stop, v_max_stop, v_max_harddecel;
    double a_max[2], a_desired[2], v_max_riding[2];
    double l, delta_max, k_p_v, k_p_delta, g;
    /* InvPendulumBicycle (parameters.py:1429-1472, :1832-1892) */
    double l_2, tau_1_squared, i_steer, c_steer, v_max_walk, delta_max_walk;
    double kx_table[5][4], ku_table[4];
    /* PlanarPointBicycle (dynamics.py:933-940) */
    double k_psi;
    /* BalancingRiderBicycle: A(v) = br_A0 + v*br_A1 + v*v*br_A2 (row-major 5x5),
     * steer-torque column br_B, pole(v) = icpt + coef*v (parameters.py:1403-1411) */
    double br_A0[25], br_A1[25], br_A2[25], br_B[5];
    double br_pole_icpt[5], br_pole_coef[5];
    /* stochastic rider behaviour (parameters.py:1398-1402, controlbehavior.py:1337-1469): the rider-behaviour
     * model = Gaussian mixture over [speed, p0_real, p1_real, p1_imag, p2_real, p2_imag] in a transformed
     * feature space; per component c: weight, mean / variance of the speed, the pole features' mean at the
     * speed's mean, their slope against the speed, and the lower Cholesky factor (row-major, packed) of their
     * conditional covariance.  Feature transforms: Yeo-Johnson lambda + standard scaler per feature (index 0
     * = speed), log shift x = sign * (exp(y) + a) per pole feature (sign 0: none). */
    int32_t br_stochastic, br_n_comp;
    int32_t br_fixed_gains, br_pad_;   /* br_fixed_gains: the gains never change (parameters `gains=`, dynamics.py:606-607) */
    double br_resample_thresh;
    uint64_t br_seed;
    double br_lam[6], br_sc_mean[6], br_sc_scale[6];
    double br_log_a[5], br_log_sign[5];
    double br_w[4], br_mu_g[4], br_var_g[4];
    double br_mu[4][5], br_slope[4][5], br_chol[4][15];
    /* payload packing */
    double q_scale;
    double q_origin[2]; /* origin (m) of the Q-format frame of the f32 payload: xq = rint((x - q_origin[0]) / q_scale) */
    int32_t traj_len;  /* int(30/t_s) = 3000, vehicle.py:159 */
    int32_t hist_len;  /* int(1/t_s)  = 100,  vehicle.py:1487 */
    int32_t hist_cap;  /* rows allocated in hist_x/hist_y (power of two > hist_len) */
    int32_t q_cap;     /* destination-queue capacity (entries per agent) */
} CsfAgentParams;

/* Per-agent state, struct of arrays of length n.  "T" is float for _f32 and
 * double for _f64 entry points; positions are always double (the f32 payload is
 * derived from them in Q format).  Pointers a model does not use may be NULL. */
typedef struct CsfAgentState {
    int64_t n;              /* array length (row stride of the 2-D fields) */
    int64_t first, count;   /* agents [first, first+count) are processed by a call */
    int64_t payload_offset; /* index of this group's agent 0 in the payload array */
    double* x;              /* [n] */
    double* y;              /* [n] */
    void* psi;              /* T [n]  yaw in (-pi, pi] */
    void* v;                /* T [n] */
    void* delta;            /* T [n]  steer angle            (twod, invpendulum, balancingrider, bicycle) */
    void* theta;            /* T [n]  roll angle             (invpendulum: theta, balancingrider: phi) */
    void* deltadot;         /* T [n]                         (balancingrider) */
    void* thetadot;         /* T [n]                         (balancingrider) */
    void* vd_default;       /* T [n]  params.v_desired_default per agent */
    int32_t* step_i;        /* [n]  Vehicle.i (ring index; wraps at traj_len for twod/invpendulum/bicycle) */
    /* destinations (Vehicle.destqueue/destpointer, vehicle.py:183-185) */
    const double* destq;    /* [n][q_cap][3]  x, y, stop */
    const int32_t* dest_len;/* [n] */
    int32_t* dest_ptr;      /* [n] */
    /* navigation state machine (vehicle.py:354-457) */
    int32_t* znav;          /* [n]  bit0 go, bit1 decelerating, bit2 arrived */
    void* znav_v0;          /* T [n]  znavparams[0..2] */
    void* znav_d0;          /* T [n] */
    void* znav_d1;          /* T [n] */
    /* history needed by the spline destination force (vehicle.py:1468-1492) */
    double* prev_x;         /* [n]  traj[0, i-1] */
    double* prev_y;         /* [n] */
    double* hist_x;         /* [hist_cap][n]  ring of past positions, row = step mod hist_cap */
    double* hist_y;
    int32_t* hist_step;     /* [n]  total steps taken (ring write index) */
    /* InvPendulumBicycle (vehicle.py:1728-1736, :1932-1950) */
    double* ip_x;           /* [5][n]  delta, deltadot, theta, thetadot, psi (unwrapped) */
    int32_t* ip_zrid;       /* [n]  bit0 riding, bit1 walking */
    int32_t* ip_delta_run;  /* [n]  trailing count of samples with |delta| < delta_max_walk */
    /* BalancingRiderBicycle / PlanarPointBicycle dynamics state (dynamics.py:305-306, :828) */
    double* dyn_x;          /* balancingrider: [5][n] phi,delta,phidot,deltadot,psi in the bike frame (unwrapped);
                             * planarpoint: [1][n] psi (unwrapped).  Positions live in x/y. */
    double* dyn_v;          /* [n] */
    double* br_gains;       /* [5][n] */
    /* stochastic rider behaviour: the agent's current pole features, the speed they were drawn at, and
     * the number of random draws it has consumed (NULL unless CsfAgentParams.br_stochastic) */
    double* br_poles;       /* [5][n] */
    double* br_vlast;       /* [n] */
    int32_t* br_draws;      /* [n] */
    int64_t* br_stream;     /* [n] key of the rider's random stream (counter-based generator: the draws of a
                             * rider depend on (seed, br_stream, draw number) only, not on grouping or sharding) */
    /* device-side status word: bit0 non-finite force/state seen, bit1 invalid nav state,
     * bit2 payload position out of Q range, bit3 degenerate spline (duplicate points),
     * bit4 no valid pole sample in 1000 draws (stochastic rider behaviour) */
    int32_t* status;        /* [1] */
    /* optional mirror in host-mapped (pinned) memory: set to 1 whenever `status` gets a bit, so that the
     * host can poll every step without a copy or a synchronisation; may be NULL */
    int32_t* status_host;   /* [1] */
} CsfAgentState;

/* ---- library ------------------------------------------------------------------ */
int csf_version(void);
const char* csf_last_error_string(void);
/* sm count of the current device (grid sizing), cached */
int csf_sm_count(void);

/* ---- K1: all-pairs repulsive force ---------------------------------------------
 * Replaces get_untracked_foes + the pair loop + np.sum of calc_forces
 * (intersection.py:690-745, :788-823, :841-843) and TwoDBicycle.calcRepulsiveForce
 * (vehicle.py:1560-1648) for one class of sources:
 *     frep[j] (+)= sum_i mask(i,j) * F(source i -> target j),  j in [0, n_tgt)
 * A coincident pair (rho == 0, which includes i == j) contributes 0.
 * `workspace` must hold csf_pair_workspace_bytes(n_src, n_tgt, sizeof scalar) bytes. */
size_t csf_pair_workspace_bytes(int64_t n_src, int64_t n_tgt, int elem_bytes);
int csf_pair_forces_f32(const void* src_xycs, int64_t n_src, const void* tgt_xycs, int64_t n_tgt,
                        const CsfFieldParams* fp, float* frep_xy, int accumulate,
                        void* workspace, size_t workspace_bytes, csf_stream_t stream);
int csf_pair_forces_f64(const void* src_xycs, int64_t n_src, const void* tgt_xycs, int64_t n_tgt,
                        const CsfFieldParams* fp, double* frep_xy, int accumulate,
                        void* workspace, size_t workspace_bytes, csf_stream_t stream);
/* Batched independent scenarios (block-diagonal interaction): agents
 * [k*group, (k+1)*group) only see each other. */
int csf_pair_forces_grouped_f32(const void* xycs, int64_t n, int32_t group, const CsfFieldParams* fp,
                                float* frep_xy, csf_stream_t stream);
int csf_pair_forces_grouped_f64(const void* xycs, int64_t n, int32_t group, const CsfFieldParams* fp,
                                double* frep_xy, csf_stream_t stream);

/* ---- K1 for sources with the v0.1 `Bicycle` elliptic field (field_kind 1) ---------------------
 * Replaces Bicycle.calcRepulsiveForce / calcPotential / updateExcentricity (vehicle.py:1054-1147)
 * inside the same masked sum.  src_e[i] = eccentricity of source i (csf_bicycle_eccentricity_*:
 * e = min((v / v_max_riding[1])^0.1, 0.7)). */
int csf_pair_forces_bicycle_f32(const void* src_xycs, const void* src_e, int64_t n_src, const void* tgt_xycs,
                                int64_t n_tgt, const CsfFieldParams* fp, float* frep_xy, int accumulate,
                                csf_stream_t stream);
int csf_pair_forces_bicycle_f64(const void* src_xycs, const void* src_e, int64_t n_src, const void* tgt_xycs,
                                int64_t n_tgt, const CsfFieldParams* fp, double* frep_xy, int accumulate,
                                csf_stream_t stream);
int csf_bicycle_eccentricity_f32(const float* v, int64_t n, double v_max, float* e, csf_stream_t stream);
int csf_bicycle_eccentricity_f64(const double* v, int64_t n, double v_max, double* e, csf_stream_t stream);

/* ---- K1, tiled variant with hierarchical culling (production path for large crowds) ------
 * Same result as csf_pair_forces_* up to the order of summation; the f32 build additionally
 * drops sources further than csf_field_cutoff_distance(fp) from the target, where every
 * contribution is below 2^-cutoff_log2 f_0 (|F| = f_0 exp(-rho q/sigma), vehicle.py:1613-1648);
 * the f64 build never truncates.  Sources are passed as a spatially sorted, tile-padded copy
 * with one bounding record per tile of 64 and per chunk of 16 tiles:
 *   csf_morton_keys_*   : int64 space-filling-curve key per payload element; (x0, y0) = lower corner
 *                         of the road users' bounding box and cell = its larger side / 65535, in
 *                         payload units (f32: integer position units, f64: metres); the caller
 *                         sorts the keys (any sort) to obtain `perm`
 *   csf_tile_sources_*  : sorted <- xycs[perm[.]] (perm NULL = identity) in the kernel's tile
 *                         layout, csf_tiled_padded_sources(n) elements; tiles:
 *                         csf_tiled_num_tiles(n) records of csf_tiled_tile_bytes(elem) bytes
 *   csf_pair_forces_tiled_* : the pair force.  `tgt_perm` (may be NULL) = visiting order of the
 *                         targets (a spatial order makes the block-level culling effective;
 *                         frep is always written in target order); `stats` (may be NULL)
 *                         accumulates the number of pair evaluations actually executed.
 *                         Work items (target block x chunk group, csf_tiled_num_items of them) are
 *                         handed out dynamically; `item_cost` (may be NULL) receives a cost measure
 *                         per item, `item_order` (may be NULL) is the order to hand them out in
 *   csf_tiled_item_order : item_order <- items by decreasing cost (<= 4096 items), to be passed to
 *                         the next launches: the heaviest items start first, the launch has no
 *                         tail.  Scheduling only: the forces do not depend on it. */
#define CSF_TILED_PREPARED 1   /* csf_tiled_prepare_* has run: block bounds and item counter are in the workspace */
#define CSF_TILED_NO_REDUCE 2  /* leave the partial sums in the workspace (csf_agent_step_fused_* reduces them) */
#define CSF_TILED_PDL 4        /* launch the pair kernel as a programmatic dependent of the preceding kernel in the stream
                                * (csf_tiled_prepare_*): its CTAs set up shared memory while that kernel drains */
int64_t csf_tiled_padded_sources(int64_t n_src);
int64_t csf_tiled_num_tiles(int64_t n_src);
int csf_tiled_tile_bytes(int elem_bytes);
size_t csf_pair_tiled_workspace_bytes(int64_t n_src, int64_t n_tgt, int elem_bytes);
double csf_field_cutoff_distance(const CsfFieldParams* fp); /* metres; INFINITY if unbounded */
/* host-only: the direction-dependent reach table the f32 kernel filters sources with, in metres:
 * reach_m[b] >= the largest distance at which |F| >= 2^-cutoff_log2 f_0 for any direction phi
 * (from the source's heading) with cos(phi) <= -1 + 2(b+1)/n_bins and any heading difference.
 * n_bins must be 64.  Returns 0, or -1 for a wrong n_bins. */
int csf_field_reach_table(const CsfFieldParams* fp, int n_bins, double* reach_m);
int csf_morton_keys_f32(const void* xycs, int64_t n, double x0, double y0, double cell, int64_t* keys,
                        csf_stream_t stream);
int csf_morton_keys_f64(const void* xycs, int64_t n, double x0, double y0, double cell, int64_t* keys,
                        csf_stream_t stream);
/* same keys with the bounding box {xmin, xmax, ymin, ymax} (payload units, double) read from DEVICE
 * memory, so that a caller that reduces the box on the device needs no host round trip */
int csf_spatial_keys_f32(const void* xycs, int64_t n, const double* box_dev, int64_t* keys, csf_stream_t stream);
int csf_spatial_keys_f64(const void* xycs, int64_t n, const double* box_dev, int64_t* keys, csf_stream_t stream);
/* The whole re-sort on the device, inside this library (no host round trip, no framework op):
 *   csf_spatial_bbox_*  : box_dev[4] <- {xmin, xmax, ymin, ymax} of the payload positions (payload units)
 *   csf_spatial_order_* : perm <- indices of xycs[0..n) sorted along the Hilbert curve over box_dev
 *                         (32-bit keys, stable radix sort); `workspace`: csf_spatial_order_workspace_bytes(n) */
size_t csf_spatial_order_workspace_bytes(int64_t n);
int csf_spatial_bbox_f32(const void* xycs, int64_t n, double* box_dev, csf_stream_t stream);
int csf_spatial_bbox_f64(const void* xycs, int64_t n, double* box_dev, csf_stream_t stream);
int csf_spatial_order_f32(const void* xycs, int64_t n, const double* box_dev, int64_t* perm, void* workspace,
                          size_t workspace_bytes, csf_stream_t stream);
int csf_spatial_order_f64(const void* xycs, int64_t n, const double* box_dev, int64_t* perm, void* workspace,
                          size_t workspace_bytes, csf_stream_t stream);
int csf_tile_sources_f32(const void* xycs, int64_t n, const int64_t* perm, void* sorted, void* tiles,
                         csf_stream_t stream);
int csf_tile_sources_f64(const void* xycs, int64_t n, const int64_t* perm, void* sorted, void* tiles,
                         csf_stream_t stream);
/* Sources with the v0.1 `Bicycle` elliptic field (fp->field_kind 1; Bicycle.calcRepulsiveForce / calcPotential /
 * updateExcentricity, vehicle.py:1054-1147) through the tiled kernel: the sorted copy carries every source's
 * heading scaled by its eccentricity e = min((speed / v_max)^0.1, 0.7) (`speed`: the sources' v column, same
 * indexing as xycs; v_max = v_max_riding[1]); csf_pair_forces_tiled_* with the same fp then evaluates that field,
 * culling with the elliptic level sets of the potential (f32: beyond 2^-cutoff_log2 p_0 / p_decay). */
int csf_tile_sources_bicycle_f32(const void* xycs, const float* speed, double v_max, int64_t n, const int64_t* perm,
                                 void* sorted, void* tiles, csf_stream_t stream);
int csf_tile_sources_bicycle_f64(const void* xycs, const double* speed, double v_max, int64_t n, const int64_t* perm,
                                 void* sorted, void* tiles, csf_stream_t stream);
int csf_pair_forces_tiled_f32(const void* sorted, const void* tiles, int64_t n_src, const void* tgt_xycs,
                              const int64_t* tgt_perm, int64_t n_tgt, const CsfFieldParams* fp,
                              float* frep_xy, int accumulate, void* workspace, size_t workspace_bytes,
                              const unsigned int* item_order, unsigned int* item_cost,
                              unsigned long long* stats, int flags, csf_stream_t stream);
int csf_pair_forces_tiled_f64(const void* sorted, const void* tiles, int64_t n_src, const void* tgt_xycs,
                              const int64_t* tgt_perm, int64_t n_tgt, const CsfFieldParams* fp,
                              double* frep_xy, int accumulate, void* workspace, size_t workspace_bytes,
                              const unsigned int* item_order, unsigned int* item_cost,
                              unsigned long long* stats, int flags, csf_stream_t stream);
int64_t csf_tiled_num_items(int64_t n_src, int64_t n_tgt, int elem_bytes);
int csf_tiled_item_order(const unsigned int* item_cost, int64_t n_items, unsigned int* item_order,
                         csf_stream_t stream);

/* ---- road-edge force ------------------------------------------------------------
 * Replaces RoadEdge.calcRepulsiveForce summed over edges (intersection.py:226-242,
 * :36-48, :81-94):  froad[j] (+)= sum_k -F_0 * r^-sigma * (vertex_k - pos_j)/r.
 * vertices: double [m][2]. */
int csf_road_forces_f32(const double* x, const double* y, int64_t n, const double* vertices, int64_t m,
                        double F_0, double sigma, float* froad_xy, int accumulate, csf_stream_t stream);
int csf_road_forces_f64(const double* x, const double* y, int64_t n, const double* vertices, int64_t m,
                        double F_0, double sigma, double* froad_xy, int accumulate, csf_stream_t stream);

/* ---- K2/K3: per-agent kernels ---------------------------------------------------
 * csf_agent_forces_*  = the per-agent part of calc_forces (intersection.py:797-799,
 *   :841-862): destination force (vehicle.py:281-299, :1416-1558, :2078-2108; mutates the
 *   destination pointer and navigation state exactly like the reference), clip of the
 *   repulsive force to |F_dest| (utils.py:56-86), sum, + road force.
 *   n_total == 1 skips the clip stage (intersection.py:813, :849-851).
 * csf_agent_advance_* = Vehicle.step / <Model>.step for every agent
 *   (vehicle.py:301-328, :1218-1272, :1386-1414, :1883-1950; dynamics.py:674-705,
 *   :1051-1079) + the mirror of positions (intersection.py:660-677), written as the
 *   next step's pair payload.
 * csf_agent_step_*    = both fused (what SocialForceIntersection.step() does per agent).
 * frep_xy / froad_xy may be NULL (treated as 0).  force_xy / fdest_xy may be NULL. */
int csf_agent_forces_f32(int model, const CsfAgentState* st, const CsfAgentParams* p, int64_t n_total,
                         const float* frep_xy, const float* froad_xy, float* force_xy, float* fdest_xy,
                         csf_stream_t stream);
int csf_agent_forces_f64(int model, const CsfAgentState* st, const CsfAgentParams* p, int64_t n_total,
                         const double* frep_xy, const double* froad_xy, double* force_xy, double* fdest_xy,
                         csf_stream_t stream);
int csf_agent_advance_f32(int model, const CsfAgentState* st, const CsfAgentParams* p,
                          const float* force_xy, void* next_xycs, csf_stream_t stream);
int csf_agent_advance_f64(int model, const CsfAgentState* st, const CsfAgentParams* p,
                          const double* force_xy, void* next_xycs, csf_stream_t stream);
int csf_agent_step_f32(int model, const CsfAgentState* st, const CsfAgentParams* p, int64_t n_total,
                       const float* frep_xy, const float* froad_xy, float* force_xy, void* next_xycs,
                       csf_stream_t stream);
int csf_agent_step_f64(int model, const CsfAgentState* st, const CsfAgentParams* p, int64_t n_total,
                       const double* frep_xy, const double* froad_xy, double* force_xy, void* next_xycs,
                       csf_stream_t stream);
/* Build the pair payload from the current state (update_road_user_positions,
 * intersection.py:660-677). */
/* Gains of every BalancingRider agent for its current speed (what BalancingRiderDynamics.__init__ computes,
 * dynamics.py:305-306, :602-615); with br_stochastic it also draws the agents' first poles. */
int csf_br_init(const CsfAgentState* st, const CsfAgentParams* p, csf_stream_t stream);
int csf_pack_xycs_f32(const CsfAgentState* st, const CsfAgentParams* p, void* xycs, csf_stream_t stream);
int csf_pack_xycs_f64(const CsfAgentState* st, const CsfAgentParams* p, void* xycs, csf_stream_t stream);
/* Generic payload packer for road users that are not stepped by this library
 * (UncontrolledVehicle sources, vehicle.py:920-987). psi: double [n]. */
int csf_pack_xypsi_f32(const double* x, const double* y, const double* psi, int64_t n, double q_scale,
                       double origin_x, double origin_y, void* xycs, csf_stream_t stream);
int csf_pack_xypsi_f64(const double* x, const double* y, const double* psi, int64_t n, double q_scale,
                       double origin_x, double origin_y, void* xycs, csf_stream_t stream);

/* ---- multi-GPU: payload exchange over NVLink peer memory --------------------------------
 * One crowd sharded by agent range over the GPUs of a node: the one exchange step
 * per simulation step -- every rank's new payload entries to every other rank -- done with
 * direct peer stores and flag words instead of a collective library call.  Each rank allocates
 * one buffer with csf_peer_alloc (cudaMalloc + IPC handle), ranks swap the handles through any
 * host channel and map them with csf_peer_open, and fill a CsfPeerComm with the device addresses
 * (layout of the buffer is the caller's: payload array, then `world` data flags, `world` read
 * flags and 4 sequence words, all zero-initialised).  Per step, in stream order:
 *   csf_peer_wait_data  -> (readers of the payload: tile build, K1) -> csf_peer_signal_read
 *   -> (K2/K3 writes the own range locally) -> csf_peer_push(first_elem, n_elem)
 * Every rank must issue the same sequence.  All three are plain kernel launches with fixed
 * arguments (CUDA-graph capturable).  seq[3] != 0 afterwards = a wait timed out (~4 s). */
#define CSF_MAX_PEERS 16
typedef struct CsfPeerComm {
    int32_t world, rank;
    void* payload[CSF_MAX_PEERS];        /* payload array of every rank (own entry: local pointer) */
    uint32_t* data_flags[CSF_MAX_PEERS]; /* [world] words in rank p's buffer; rank r writes word r */
    uint32_t* read_flags[CSF_MAX_PEERS];
    uint32_t* seq;                       /* local: push count, read count, block counter, status */
    int32_t* status_host;                /* optional host-mapped mirror of seq[3] != 0 (may be NULL) */
} CsfPeerComm;
/* What the fused step kernels fold in (csf_agent_step_fused_*): the fixed-order reduction of the tiled pair
 * kernel's partial sums (launched with CSF_TILED_NO_REDUCE) and this rank's side of the payload exchange. */
typedef struct CsfStepFusion {
    const void* partial;     /* T [n_groups][partial_stride][2]: workspace + csf_tiled_partial_offset(); NULL: none */
    int64_t partial_stride;  /* n_tgt of the pair call */
    int64_t partial_offset;  /* index of the group's agent 0 among the pair call's targets */
    int32_t n_groups;        /* csf_tiled_num_groups() */
    int32_t pdl;             /* 1: launch as a programmatic dependent of the preceding kernel in the stream (the pair
                              * kernel): the part of the step that does not need the pair forces overlaps its tail */
    double f0;               /* field strength f_0 (the partial sums are per unit f_0) */
    CsfPeerComm comm;        /* comm.world <= 1: no exchange */
} CsfStepFusion;
/* csf_agent_step_* with the reduction of the pair kernel's partial sums and the payload exchange folded in:
 * one launch instead of reduce + signal + step + push. */
int csf_agent_step_fused_f32(int model, const CsfAgentState* st, const CsfAgentParams* p, int64_t n_total,
                             const CsfStepFusion* fusion, const float* froad, float* force, void* next_xycs,
                             csf_stream_t stream);
int csf_agent_step_fused_f64(int model, const CsfAgentState* st, const CsfAgentParams* p, int64_t n_total,
                             const CsfStepFusion* fusion, const double* froad, double* force, void* next_xycs,
                             csf_stream_t stream);
/* The step's first kernel: tile build (csf_tile_sources_*) + target-block bounds + item-counter reset in one
 * launch; with `comm` (may be NULL) it first waits for the peers' payload pushes (csf_peer_wait_data).
 * Follow with csf_pair_forces_tiled_*(..., flags | CSF_TILED_PREPARED). */
int csf_tiled_prepare_f32(const void* xycs, int64_t n_src, const int64_t* perm, void* sorted, void* tiles,
                          const void* tgt_xycs, const int64_t* tgt_perm, int64_t n_tgt, void* workspace,
                          size_t workspace_bytes, const CsfPeerComm* comm, csf_stream_t stream);
int csf_tiled_prepare_f64(const void* xycs, int64_t n_src, const int64_t* perm, void* sorted, void* tiles,
                          const void* tgt_xycs, const int64_t* tgt_perm, int64_t n_tgt, void* workspace,
                          size_t workspace_bytes, const CsfPeerComm* comm, csf_stream_t stream);
int csf_tiled_num_groups(int64_t n_src, int64_t n_tgt, int elem_bytes);
size_t csf_tiled_partial_offset(int64_t n_src, int64_t n_tgt, int elem_bytes);
int csf_peer_handle_bytes(void);
int csf_peer_alloc(size_t bytes, void** devptr, void* ipc_handle_out);
int csf_peer_open(const void* ipc_handle, void** devptr);
int csf_peer_close(void* devptr);
int csf_peer_free(void* devptr);
int csf_peer_wait_data(const CsfPeerComm* comm, csf_stream_t stream);
int csf_peer_signal_read(const CsfPeerComm* comm, csf_stream_t stream);
int csf_peer_push(const CsfPeerComm* comm, int64_t first_elem, int64_t n_elem, int elem_bytes,
                  csf_stream_t stream);

/* ---- trajectory stream and host bridge -------------------------------------------
 * Replaces the per-vehicle history writes of the reference (vehicle.py:320-325 `traj[:, i] = s`,
 * :1407-1413, `trajF`) and the per-vehicle traci.vehicle.moveToXY loop (intersection.py:660-688):
 *   csf_copy_segments : up to CSF_MAX_COPY_SEGMENTS contiguous device-to-device copies in one launch (the
 *                       state columns of every model group + the total forces -> one slot of a device ring
 *                       that the host drains a chunk of steps at a time); lengths are multiples of 4 bytes;
 *   csf_sumo_pose_*   : out[i] = {x, y, SUMO angle in degrees} (utils.py:119-121 angleSFMtoSUMO) for a
 *                       batched moveToXY after ONE device-to-host copy per step. */
#define CSF_MAX_COPY_SEGMENTS 16
typedef struct CsfCopySegment {
    const void* src;
    void* dst;
    int64_t bytes;
} CsfCopySegment;
typedef struct CsfCopySegments {
    int32_t n;
    int32_t pad_;
    CsfCopySegment seg[CSF_MAX_COPY_SEGMENTS];
} CsfCopySegments;
int csf_copy_segments(const CsfCopySegments* segs, csf_stream_t stream);
/* Road-user churn (intersection.py:458-539 add_road_user, :576-634 remove_road_user[s_by_id]): every per-agent
 * array of a model group re-laid in one launch.  A segment views an array as (outer, n, inner_bytes):
 *   dst[o][dst_off + j][:] = src[o][idx ? idx[j] : j][:]   for o < outer, j < count
 * (idx: device array of source agents -- select / compact; NULL -- append behind dst_off). */
#define CSF_MAX_GATHER_SEGMENTS 48
typedef struct CsfGatherSegment {
    const void* src;
    void* dst;
    const int64_t* idx;   /* [count] or NULL */
    int64_t outer, n_src, n_dst, dst_off, count, inner_bytes;
    int64_t dst_inner_bytes; /* 0: as inner_bytes; larger: the element is copied into the head of a wider destination
                              * element (destination queues of different capacity), the rest is left untouched */
} CsfGatherSegment;
typedef struct CsfGatherSegments {
    int32_t n;
    int32_t pad_;
    CsfGatherSegment seg[CSF_MAX_GATHER_SEGMENTS];
} CsfGatherSegments;
int csf_gather_segments(const CsfGatherSegments* segs, csf_stream_t stream);
int csf_sumo_pose_f32(const double* x, const double* y, const void* psi, int64_t n, double* out, csf_stream_t stream);
int csf_sumo_pose_f64(const double* x, const double* y, const void* psi, int64_t n, double* out, csf_stream_t stream);

/* ---- measurement helper: sustained FP32 FFMA throughput of this device ------------
 * Runs `iters` dependent-chain FFMA rounds on every SM; returns flops executed
 * (2 per FFMA) through *flops; the caller times it with CUDA events. */
int csf_ffma_peak(int64_t iters, float* sink, double* flops, csf_stream_t stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* CSF_B200_H */
