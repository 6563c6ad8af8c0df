// K2/K3: per-agent destination force, force assembly, rider control and bicycle
// dynamics for every model class -- one thread per agent, state in registers.
//
// Reference (src/cyclistsocialforce/):
//   Vehicle.updateDestination        vehicle.py:545-594
//   Vehicle.updateNavState           vehicle.py:354-457
//   TwoDBicycle.calcDestinationForce vehicle.py:1416-1558   (FITPACK splprep/splev restated)
//   Bicycle.calcDestinationForceField / calc_direct_approach_dest_force  :1150-1187, :2078-2108
//   calc_forces (per-agent tail)     intersection.py:841-862,  utils.limitMagnitude utils.py:56-86
//   Bicycle.control / move           vehicle.py:1218-1272
//   TwoDBicycle.step                 vehicle.py:1386-1414
//   InvPendulumBicycle.step*         vehicle.py:1738-1950, parameters.py:1832-1892
//   BalancingRiderDynamics.step      dynamics.py:602-705,  from_pole_placement :1167-1227
//   PlanarPointDynamics.step         dynamics.py:996-1079
#include "csf_common.cuh"
#include "csf_peer.cuh"
#include <string.h>

namespace {

enum { MODE_FORCES = 0, MODE_ADVANCE = 1, MODE_STEP = 2 };

// ----------------------------------------------------------------------------------------
// per-agent register image
// ----------------------------------------------------------------------------------------
constexpr int kQW = 6;  // destination-queue window held in registers: entries ptr0 .. ptr0 + 5
template <typename T> struct Agent {
    double x, y;
    T psi, v, delta, theta;
    T vd_def;
    int i;            // Vehicle.i
    int ptr, qlen;    // destination pointer / queue length
    int znav;         // bit0 go, bit1 decel, bit2 arrived
    T z_v0, z_d0, z_d1;
    const double* q;  // this agent's destination queue [q_cap][3]
    // The queue entries the step can touch, fetched with independent loads when the kernel starts,
    // relative to the position at that time (x0, y0): entry min(ptr0 + j, qlen - 1) for j < kQW.
    // updateDestination advances the pointer by at most 2 per call and the spline looks 3 entries
    // ahead, so only a step with three pointer updates in a row (spline fall-back) leaves the window;
    // it then reads the queue itself.
    double x0, y0;
    int ptr0;
    T wx[kQW], wy[kQW];
    unsigned wstop;   // bit j: stop flag of window entry j
    int flags;        // status bits raised by this agent
};

template <typename T> __device__ __forceinline__ void load_window(Agent<T>& a) {
    a.x0 = a.x;
    a.y0 = a.y;
    a.ptr0 = a.ptr;
    a.wstop = 0;
    double qx[kQW], qy[kQW], qs[kQW];
#pragma unroll
    for (int j = 0; j < kQW; ++j) {
        const int idx = min(a.ptr + j, a.qlen - 1);
        qx[j] = a.q[idx * 3 + 0];
        qy[j] = a.q[idx * 3 + 1];
        qs[j] = a.q[idx * 3 + 2];
    }
#pragma unroll
    for (int j = 0; j < kQW; ++j) {
        a.wx[j] = (T)(qx[j] - a.x);
        a.wy[j] = (T)(qy[j] - a.y);
        a.wstop |= (qs[j] != 0.0) ? (1u << j) : 0u;
    }
}
// queue entry idx relative to (x0, y0), and its stop flag
template <typename T> __device__ __forceinline__ void queue_entry(const Agent<T>& a, int idx, T& rx, T& ry, bool& stop) {
    const int j = min(idx, a.qlen - 1) - a.ptr0;
    if (j >= 0 && j < kQW) {
        rx = a.wx[0];
        ry = a.wy[0];
#pragma unroll
        for (int u = 1; u < kQW; ++u) {
            rx = (j == u) ? a.wx[u] : rx;
            ry = (j == u) ? a.wy[u] : ry;
        }
        stop = (a.wstop >> j) & 1u;
    } else {
        const int k = min(idx, a.qlen - 1);
        rx = (T)(a.q[k * 3 + 0] - a.x0);
        ry = (T)(a.q[k * 3 + 1] - a.y0);
        stop = a.q[k * 3 + 2] != 0.0;
    }
}

// distance to queue entry idx (from the position the step started at; it moves only in K3's last lines)
template <typename T> __device__ __forceinline__ T dist_to(const Agent<T>& a, int idx) {
    T dx, dy;
    bool st;
    queue_entry(a, idx, dx, dy, st);
    return qsqrt(dx * dx + dy * dy);
}

// Vehicle.updateDestination, vehicle.py:545-594
template <typename T> __device__ __forceinline__ void update_destination(Agent<T>& a, const CsfAgentParams& p) {
    const T dnext = dist_to(a, a.ptr);
    if (a.znav & 6) return;
    if (dnext <= (T)p.d_arrived_inter) a.ptr = min(a.ptr + 1, a.qlen - 1);
    if (a.ptr < a.qlen - 1) {
        const T dnn = dist_to(a, a.ptr + 1);
        if (dnn < dnext) a.ptr += 1;
    }
}

// Vehicle.updateNavState, vehicle.py:354-457 -> v_d ; *ddest_out = distance to destination
template <typename T> __device__ __forceinline__ T update_nav_state(Agent<T>& a, const CsfAgentParams& p, bool stop, T* ddest_out) {
    const T k = (T)1.5;
    const bool z0 = a.znav & 1, z1 = a.znav & 2, z2 = a.znav & 4;
    const T vhd = (T)p.v_max_harddecel;
    T d0, d1;
    if (z0) {
        d0 = (T)0.5 * (vhd * vhd - a.v * a.v) / (T)p.a_desired[0];
        d1 = (T)0.5 * -(vhd * vhd) / (T)p.a_max[0];
    } else {
        d0 = a.z_d0;
        d1 = a.z_d1;
    }
    const T dd = dist_to(a, a.ptr);
    const bool x0 = stop, x1 = dd <= k * (d0 + d1), x2 = dd <= (T)p.d_arrived_stop, x3 = a.v <= (T)p.v_max_stop;
    const bool n0 = !x0 || (x0 && !x1 && ((z0 && !x2) || z1));
    const bool n1 = x0 && ((z0 && ((!x2 && x1) || (x2 && !x3))) || (z1 && x1 && (!x2 || !x3)));
    const bool n2 = x0 && (((z0 || z1) && x2 && x3) || z2);
    a.znav = (n0 ? 1 : 0) | (n1 ? 2 : 0) | (n2 ? 4 : 0);
    if (z0 && n1) {
        a.z_v0 = a.v;
        a.z_d0 = d0;
        a.z_d1 = d1;
    }
    *ddest_out = dd;
    T vd;
    if (n0) vd = a.vd_def;
    else if (n1) {
        if (dd < k * a.z_d1) vd = vhd / a.z_d1 * dd * (T)1 / k;
        else vd = (a.z_v0 - vhd) / a.z_d0 * (dd - a.z_d1) * (T)1 / k + vhd;
    } else if (n2) vd = (T)0;
    else {
        vd = (T)0;
        a.flags |= 2;  // reference raises "Invalid navigation state"
    }
    return vd;
}

// Bicycle.calcDestinationForceField (vehicle.py:1168-1187) == calc_direct_approach_dest_force (:2096-2108)
template <typename T> __device__ __forceinline__ void dest_force_direct(Agent<T>& a, const CsfAgentParams& p, T& fx, T& fy) {
    update_destination(a, p);
    T rx, ry;
    bool stop;
    queue_entry(a, a.ptr, rx, ry, stop);
    T dd;
    const T vd = update_nav_state(a, p, stop, &dd);
    if (dd > (T)0) {
        fx = -vd * (-rx) / dd;
        fy = -vd * (-ry) / dd;
    } else {
        fx = (T)0;
        fy = (T)0;
    }
}

// ----------------------------------------------------------------------------------------
// interpolating parametric cubic B-spline through M in {4,5,6} points
// (scipy.interpolate.splprep(s=0)/splev == FITPACK parcur/fppara/splev/splder)
//
// M is a template parameter: every loop below unrolls with static indices, the knot vector
// 0,0,0,0, u_2 .. u_{M-3}, 1,1,1,1 keeps its constant entries as constants, and the collocation system
// is solved for the M - 2 interior coefficients only (c_0 and c_{M-1} are the end points): everything
// lives in registers.  The callers switch on the run-time number of points.
// ----------------------------------------------------------------------------------------
template <typename T, int M> struct SplineM {
    T kn[M + 4];      // knots
    T cx[M], cy[M];   // B-spline coefficients
};
// Cox-de Boor on the six knots K[0..5] = t_{l-2} .. t_{l+3} around interval l: cubic N3[0..3]
// (B_{l-3..l}), quadratic N2[0..2], linear N1[0..1]
template <typename T> __device__ __forceinline__ void basis_local(const T* K, T u, T* N3, T* N2, T* N1) {
    T left[4], right[4], N[4];
#pragma unroll
    for (int j = 1; j <= 3; ++j) {
        left[j] = u - K[3 - j];
        right[j] = K[2 + j] - u;
    }
    N[0] = (T)1;
#pragma unroll
    for (int j = 1; j <= 3; ++j) {
        T saved = (T)0;
#pragma unroll
        for (int r = 0; r < j; ++r) {
            const T temp = qdiv(N[r], right[r + 1] + left[j - r]);
            N[r] = saved + right[r + 1] * temp;
            saved = left[j - r] * temp;
        }
        N[j] = saved;
        if (j == 1) { N1[0] = N[0]; N1[1] = N[1]; }
        if (j == 2) { N2[0] = N[0]; N2[1] = N[1]; N2[2] = N[2]; }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) N3[j] = N[j];
}
// Fit: px,py = control points (relative coordinates); returns false on duplicate points.
template <typename T, int M> __device__ __forceinline__ bool spline_fit(SplineM<T, M>& s, const T* px, const T* py) {
    T u[M];
    u[0] = (T)0;
    bool ok = true;
#pragma unroll
    for (int k = 1; k < M; ++k) {                      // chord-length parameters
        const T dx = px[k] - px[k - 1], dy = py[k] - py[k - 1];
        const T d = qsqrt(dx * dx + dy * dy);
        ok = ok && (d > (T)0);
        u[k] = u[k - 1] + d;
    }
    const T tot = u[M - 1];
    const T itot = qdiv((T)1, tot);
#pragma unroll
    for (int k = 1; k < M - 1; ++k) u[k] = u[k] * itot;
    u[M - 1] = (T)1;
#pragma unroll
    for (int i = 0; i < M + 4; ++i) s.kn[i] = i < 4 ? (T)0 : (i >= M ? (T)1 : u[i - 2]);
    if (!ok) return false;
    // collocation at u_1 .. u_{M-2}; u_i lies in knot interval l = min(i + 2, M - 1), where the basis
    // functions B_{l-3} .. B_l are the non-zero ones; columns 0 and M - 1 go to the right-hand side
    constexpr int NI = M - 2;
    T A[NI][NI], bx[NI], by[NI];
#pragma unroll
    for (int i = 1; i <= NI; ++i) {
        const int l = i + 2 < M - 1 ? i + 2 : M - 1;
        T N3[4], N2[3], N1[2];
        basis_local(&s.kn[l - 2], u[i], N3, N2, N1);
        T rx = px[i], ry = py[i];
#pragma unroll
        for (int j = 0; j < NI; ++j) A[i - 1][j] = (T)0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int col = l - 3 + r;
            if (col == 0) { rx -= N3[r] * px[0]; ry -= N3[r] * py[0]; }
            else if (col == M - 1) { rx -= N3[r] * px[M - 1]; ry -= N3[r] * py[M - 1]; }
            else A[i - 1][col - 1] = N3[r];
        }
        bx[i - 1] = rx;
        by[i - 1] = ry;
    }
    // Gaussian elimination without pivoting (B-spline collocation matrices are totally positive)
#pragma unroll
    for (int k = 0; k < NI; ++k) {
        const T inv = qdiv((T)1, A[k][k]);
#pragma unroll
        for (int i = k + 1; i < NI; ++i) {
            const T f = A[i][k] * inv;
#pragma unroll
            for (int j = k + 1; j < NI; ++j) A[i][j] -= f * A[k][j];
            bx[i] -= f * bx[k];
            by[i] -= f * by[k];
        }
    }
    s.cx[0] = px[0]; s.cy[0] = py[0];
    s.cx[M - 1] = px[M - 1]; s.cy[M - 1] = py[M - 1];
#pragma unroll
    for (int k = NI - 1; k >= 0; --k) {
        T sx = bx[k], sy = by[k];
#pragma unroll
        for (int j = k + 1; j < NI; ++j) {
            sx -= A[k][j] * s.cx[j + 1];
            sy -= A[k][j] * s.cy[j + 1];
        }
        const T inv = qdiv((T)1, A[k][k]);
        s.cx[k + 1] = sx * inv;
        s.cy[k + 1] = sy * inv;
    }
    return true;
}
// S(u) and optionally S'(u), S''(u).  The knot interval of u is a run-time value in [3, M - 1]: the six
// knots and four coefficients around it are picked with selects over that (static) range.
template <typename T, int M>
__device__ __forceinline__ void spline_eval(const SplineM<T, M>& s, T u, bool der, T& x, T& y, T& dx, T& dy, T& ddx,
                                            T& ddy) {
    int l = 3;  // FITPACK splev: advance while u >= t[l+1] and l != M-1
#pragma unroll
    for (int ll = 4; ll <= M - 1; ++ll)
        if (l == ll - 1 && u >= s.kn[ll]) l = ll;
    T K[6], cxl[4], cyl[4];
#pragma unroll
    for (int j = 0; j < 6; ++j) K[j] = s.kn[1 + j];
#pragma unroll
    for (int r = 0; r < 4; ++r) { cxl[r] = s.cx[r]; cyl[r] = s.cy[r]; }
#pragma unroll
    for (int ll = 4; ll <= M - 1; ++ll) {
#pragma unroll
        for (int j = 0; j < 6; ++j) K[j] = (l == ll) ? s.kn[ll - 2 + j] : K[j];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            cxl[r] = (l == ll) ? s.cx[ll - 3 + r] : cxl[r];
            cyl[r] = (l == ll) ? s.cy[ll - 3 + r] : cyl[r];
        }
    }
    T N3[4], N2[3], N1[2];
    basis_local(K, u, N3, N2, N1);
    x = y = (T)0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        x += N3[r] * cxl[r];
        y += N3[r] * cyl[r];
    }
    if (!der) return;
    T d1x[3], d1y[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const T w = qdiv((T)3, K[3 + r] - K[r]);
        d1x[r] = w * (cxl[r + 1] - cxl[r]);
        d1y[r] = w * (cyl[r + 1] - cyl[r]);
    }
    dx = N2[0] * d1x[0] + N2[1] * d1x[1] + N2[2] * d1x[2];
    dy = N2[0] * d1y[0] + N2[1] * d1y[1] + N2[2] * d1y[2];
    const T w0 = qdiv((T)2, K[3] - K[1]), w1 = qdiv((T)2, K[4] - K[2]);
    ddx = N1[0] * (w0 * (d1x[1] - d1x[0])) + N1[1] * (w1 * (d1x[2] - d1x[1]));
    ddy = N1[0] * (w0 * (d1y[1] - d1y[0])) + N1[1] * (w1 * (d1y[2] - d1y[1]));
}

// The spline part of TwoDBicycle.calcDestinationForce (vehicle.py:1496-1553) for M control points:
// returns 0 = force set, 1 = duplicate points (FITPACK would raise), 2 = look-ahead sample beyond the curve
template <typename T, int M>
__device__ __forceinline__ int spline_force(const T* px, const T* py, int cur, bool last, bool stop, T vd, T g,
                                            T& fx, T& fy) {
    SplineM<T, M> sp;
    if (!spline_fit<T, M>(sp, px, py)) return 1;
    const T step = (T)1 / (T)19;
    int i_s = 1;
    if (last) {  // argmin over the 20 samples  :1516-1520
        T best = (T)0;
        for (int j = 0; j < 20; ++j) {
            const T u = (j == 19) ? (T)1 : (T)j * step;
            T x, y, t0, t1, t2, t3;
            spline_eval<T, M>(sp, u, false, x, y, t0, t1, t2, t3);
            const T ex = x - px[cur], ey = y - py[cur];
            const T d2 = ex * ex + ey * ey;
            if (j == 0 || d2 < best) { best = d2; i_s = j; }
        }
    }
    const int i_p = i_s + (stop ? 5 : 3);  // :1523-1526
    if (i_p >= 20) return 2;
    const T us = (i_s == 19) ? (T)1 : (T)i_s * step;
    const T up = (i_p == 19) ? (T)1 : (T)i_p * step;
    T sx, sy, dx, dy, ddx, ddy;
    spline_eval<T, M>(sp, us, true, sx, sy, dx, dy, ddx, ddy);
    T qx, qy, t0, t1, t2, t3;
    spline_eval<T, M>(sp, up, false, qx, qy, t0, t1, t2, t3);
    const T sp1 = qsqrt(dx * dx + dy * dy);
    const T R = qdiv(sp1 * sp1 * sp1, fabs(dx * ddy - dy * ddx));  // :1532-1537
    const T thetacomf = (T)(10.0 * (CSF_TWO_PI / 360.0));
    T v = fmax((T)2.5, qsqrt(thetacomf * g * R));
    v = fmin(v, vd);
    const T ex = qx - sx, ey = qy - sy;
    const T temp = qdiv(v, qsqrt(ex * ex + ey * ey));
    fx = temp * ex;
    fy = temp * ey;
    return 0;
}

// TwoDBicycle.calcDestinationForce, vehicle.py:1443-1558
template <typename T>
__device__ __forceinline__ void dest_force_twod(Agent<T>& a, const CsfAgentParams& p, const CsfAgentState& st, int64_t k, double pv_x,
                                double pv_y, T& fx, T& fy) {
    update_destination(a, p);
    T drx, dry;
    bool stop;
    queue_entry(a, a.ptr, drx, dry, stop);
    T dd;
    const T vd = update_nav_state(a, p, stop, &dd);
    if (a.i == 0) {  // :1455-1458
        T sn, cs;
        sincosT(a.psi, &sn, &cs);
        fx = vd * cs;
        fy = vd * sn;
        return;
    }
    if (a.znav & 4) {  // :1461-1462
        fx = fy = (T)0;
        return;
    }
    const bool last = a.ptr + 1 >= a.qlen;
    // control points relative to the current position (index of the current position: 1 or 2)
    T px[6], py[6];
    int m, cur;
    const T pvx = (T)(pv_x - a.x), pvy = (T)(pv_y - a.y);
    if (!last) {  // [traj[i-1], traj[i], destqueue[ptr : ptr+4]]  :1468-1479
        px[0] = pvx; py[0] = pvy;
        px[1] = (T)0; py[1] = (T)0;
        const int nd = min(4, a.qlen - a.ptr);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            bool sj;
            queue_entry(a, a.ptr + j, px[2 + j], py[2 + j], sj);
        }
        m = 2 + nd;
        cur = 1;
    } else {  // [traj[max(0, i-100)], traj[i-1], traj[i], dest]  :1486-1492
        const int back = min(a.i, p.hist_len);
        const int hs = st.hist_step[k];      // (the last-destination branch only)
        const int row = (hs - back) & (p.hist_cap - 1);
        px[0] = (T)(st.hist_x[(size_t)row * st.n + k] - a.x);
        py[0] = (T)(st.hist_y[(size_t)row * st.n + k] - a.y);
        px[1] = pvx; py[1] = pvy;
        px[2] = (T)0; py[2] = (T)0;
        px[3] = drx; py[3] = dry;
        px[4] = px[5] = py[4] = py[5] = (T)0;
        m = 4;
        cur = 2;
    }
    int rc;
    if (m == 6) rc = spline_force<T, 6>(px, py, cur, last, stop, vd, (T)p.g, fx, fy);
    else if (m == 5) rc = spline_force<T, 5>(px, py, cur, last, stop, vd, (T)p.g, fx, fy);
    else rc = spline_force<T, 4>(px, py, cur, last, stop, vd, (T)p.g, fx, fy);
    if (rc != 0) {
        if (rc == 1) a.flags |= 8;  // reference: FITPACK ValueError (duplicate points); here: direct approach
        dest_force_direct(a, p, fx, fy);  // (rc == 2: vehicle.py:1556)
    }
}

template <typename T, int MODEL>
__device__ __forceinline__ void destination_force(Agent<T>& a, const CsfAgentParams& p, const CsfAgentState& st, int64_t k, double pv_x,
                                  double pv_y, T& fx, T& fy) {
    if (MODEL == CSF_MODEL_TWOD || MODEL == CSF_MODEL_INVPENDULUM) dest_force_twod(a, p, st, k, pv_x, pv_y, fx, fy);
    else if (MODEL == CSF_MODEL_BICYCLE) dest_force_direct(a, p, fx, fy);              // vehicle.py:1189-1194
    else if (MODEL == CSF_MODEL_BALANCINGRIDER) {
        update_destination(a, p);                                                      // vehicle.py:295-297
        dest_force_direct(a, p, fx, fy);
    } else {                                                                           // planarpoint
        update_destination(a, p);                                                      // vehicle.py:295-297
        dest_force_twod(a, p, st, k, pv_x, pv_y, fx, fy);                              // vehicle.py:2025
    }
}

// ----------------------------------------------------------------------------------------
// Bicycle.control + move, vehicle.py:1218-1272
// ----------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ void control_move(Agent<T>& a, const CsfAgentParams& p, T Fx, T Fy) {
    const T th = atan2(Fy, Fx);
    T vF = sqrt(Fx * Fx + Fy * Fy);
    const T dd = dist_to(a, a.ptr);
    if (dd < (T)3 && (a.ptr + 1 >= a.qlen)) vF = (vF / (T)3) * dd;
    const T target = angle_difference(a.psi, th);
    const T ddelta = angle_difference(a.delta, target);
    T acc = (T)p.k_p_v * (vF - a.v);
    const T od = (T)p.k_p_delta * ddelta;
    const T ts = (T)p.t_s;
    acc = clampT(acc, (T)p.a_max[0], (T)p.a_max[1]);
    T delta = limit_angle(a.delta + ts * od);
    T v = a.v + ts * acc;
    delta = clampT(delta, (T)-p.delta_max, (T)p.delta_max);
    v = clampT(v, (T)p.v_max_riding[0], (T)p.v_max_riding[1]);
    const T psi = limit_angle(a.psi + ts * v * tan(delta) / (T)p.l);
    T sn, cs;
    sincosT(psi, &sn, &cs);
    a.y += (double)(ts * v * sn);
    a.x += (double)(ts * v * cs);
    a.psi = psi;
    a.v = v;
    a.delta = delta;
}

// ----------------------------------------------------------------------------------------
// small dense double-precision linear algebra (per-agent; every loop unrolled, so the arrays live in
// registers -- no local memory)
// ----------------------------------------------------------------------------------------
// solve A x = b (5x5, partial pivoting by predicated row swaps), A and b destroyed
__device__ __forceinline__ void solve5(double (&A)[5][5], double (&b)[5], double (&x)[5]) {
    constexpr int N = 5;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        // bring the largest |A[i][k]|, i >= k, to row k: compare-and-swap with every later row (the first
        // of several equal candidates stays, as in the reference's LAPACK pivoting)
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            const bool sw = fabs(A[i][k]) > fabs(A[k][k]);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const double u = A[k][j], w = A[i][j];
                A[k][j] = sw ? w : u;
                A[i][j] = sw ? u : w;
            }
            const double u = b[k], w = b[i];
            b[k] = sw ? w : u;
            b[i] = sw ? u : w;
        }
        const double inv = 1.0 / A[k][k];
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            const double f = A[i][k] * inv;
#pragma unroll
            for (int j = k + 1; j < N; ++j) A[i][j] = fma(-f, A[k][j], A[i][j]);
            b[i] = fma(-f, b[k], b[i]);
        }
    }
#pragma unroll
    for (int k = N - 1; k >= 0; --k) {
        double t = b[k];
#pragma unroll
        for (int j = k + 1; j < N; ++j) t = fma(-A[k][j], x[j], t);
        x[k] = t / A[k][k];
    }
}

// ----------------------------------------------------------------------------------------
// InvPendulumBicycle, vehicle.py:1738-1950 (dynamics in double in both builds)
// ----------------------------------------------------------------------------------------
__constant__ double kInvK[25] = {0.0, 1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10, 1.0 / 11,
                                 1.0 / 12, 1.0 / 13, 1.0 / 14, 1.0 / 15, 1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19, 1.0 / 20, 1.0 / 21,
                                 1.0 / 22, 1.0 / 23, 1.0 / 24};
// floor(log2 |a|) of a finite, normal, non-zero double from its exponent field (0 for a == 0), and 2^e as a
// double built the same way (|e| < 1022): scaling by a power of two is then one exact multiplication --
// ilogb / scalbn are library calls with dozens of instructions each
__device__ __forceinline__ int exp2_of(double a) {
    const int ex = (__double2hiint(a) >> 20) & 0x7ff;
    return ex == 0 ? 0 : ex - 1023;
}
__device__ __forceinline__ double pow2(int e) { return __hiloint2double((1023 + e) << 20, 0); }

// x+ = [I 0] exp(M) [x; psi_d],  M = [[A_c t_s, B_c t_s], [0, 0]]  -- what ct.forced_response returns for a
// constant input over one sample (vehicle.py:1835-1842).  The reference forms the full matrix exponential
// (scipy's Pade-13 scaling-and-squaring); only its ACTION on one vector is needed, and M has 12 non-zero
// entries.  |M|_1 is 50...400 (steer-torque gains over a small steering inertia) although its spectral
// radius is below 1: M is badly scaled, not large.  A diagonal similarity by powers of two -- exact in
// floating point -- brings |D^-1 M D|_1 to 2...4; then  exp(M) y = D (T_24(D^-1 M D / n))^n D^-1 y  with the
// degree-24 Taylor polynomial in Horner form and n = 1, 2, 4 ... sub-steps so that the norm per sub-step is
// at most 2 (truncation 2^25/25! < 3e-18).  ~16 FMAs per term, everything in registers; agrees with
// scipy.linalg.expm to 1e-15 relative (tests: invpendulum crowds against the oracle).
__device__ __forceinline__ void invpend_yaw_step(const CsfAgentParams& p, double v, double psi_d, double (&x)[5]) {
    // open loop (:1738-1768) with time-varying K, K tau_2, tau_3 (parameters.py:1850-1855)
    const double l = p.l;
    const double K_tau_2 = (v * p.l_2) / (p.g * l), K = (v * v) / (p.g * l), tau_3 = l / v;
    const double iv = 1.0 / v;
    const double vd[4] = {1.0, iv, iv * iv, iv * iv * iv};
    double kx[5], ku = 0.0;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        kx[r] = 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) kx[r] += p.kx_table[r][c] * vd[c];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) ku += p.ku_table[c] * vd[c];
    const double bI = 1.0 / p.i_steer;
    const double ts = p.t_s;
    // A_c = A - B K_x ; B_c = K_u B  (rows scaled by t_s)
    double m01 = ts, m23 = ts;
    double m1[6];
#pragma unroll
    for (int c = 0; c < 5; ++c) m1[c] = -bI * kx[c] * ts;
    m1[1] += (-p.c_steer * bI) * ts;
    m1[5] = ku * bI * ts;
    double m30 = -K / p.tau_1_squared * ts, m31 = -K_tau_2 / p.tau_1_squared * ts, m32 = 1.0 / p.tau_1_squared * ts;
    double m40 = 1.0 / tau_3 * ts;
    // balancing exponents e_i (d_i = 2^e_i, d_0 = 1): equalise the magnitudes of the entry pairs (0,1)/(1,0),
    // (3,1)/(1,3), (3,2)/(2,3), (4,0)/(1,4); the input column gets magnitude ~1
    const int e1 = (exp2_of(m1[0]) - exp2_of(m01)) >> 1;
    const int e3 = e1 + ((exp2_of(m31) - exp2_of(m1[3])) >> 1);
    const int e2 = e3 - ((exp2_of(m32) - exp2_of(m23)) >> 1);
    const int e4 = (e1 + exp2_of(m40) - exp2_of(m1[4])) >> 1;
    const int e5 = e1 - exp2_of(m1[5]);
    const int e[6] = {0, e1, e2, e3, e4, e5};
    m01 *= pow2(e[1] - e[0]);
    m23 *= pow2(e[3] - e[2]);
#pragma unroll
    for (int c = 0; c < 6; ++c) m1[c] *= pow2(e[c] - e[1]);
    m30 *= pow2(e[0] - e[3]);
    m31 *= pow2(e[1] - e[3]);
    m32 *= pow2(e[2] - e[3]);
    m40 *= pow2(e[0] - e[4]);
    // 1-norm of the balanced matrix -> sub-steps
    double nrm = fabs(m1[0]) + fabs(m30) + fabs(m40);
    nrm = fmax(nrm, fabs(m01) + fabs(m1[1]) + fabs(m31));
    nrm = fmax(nrm, fabs(m1[2]) + fabs(m32));
    nrm = fmax(nrm, fabs(m1[3]) + fabs(m23));
    nrm = fmax(nrm, fmax(fabs(m1[4]), fabs(m1[5])));
    int ns = 1, sh = 0;
    while (nrm > 2.0 * ns && sh < 20) { ns <<= 1; ++sh; }
    const double isub = pow2(-sh);
    m01 *= isub;
    m23 *= isub;
#pragma unroll
    for (int c = 0; c < 6; ++c) m1[c] *= isub;
    m30 *= isub;
    m31 *= isub;
    m32 *= isub;
    m40 *= isub;
    double y[5];
#pragma unroll
    for (int r = 0; r < 5; ++r) y[r] = x[r] * pow2(-e[r]);
    const double c15 = m1[5] * (psi_d * pow2(-e[5]));       // the input is constant over the step
    for (int q = 0; q < ns; ++q) {
        double r0 = y[0], r1 = y[1], r2 = y[2], r3 = y[3], r4 = y[4];
#pragma unroll 4
        for (int kk = 24; kk >= 1; --kk) {
            const double ik = kInvK[kk];                     // (rolled: the unrolled loop is 400 instructions that
                                                             //  every SM fetches once per launch)
            const double t0 = m01 * r1;
            const double t1 = fma(m1[0], r0, fma(m1[1], r1, fma(m1[2], r2, fma(m1[3], r3, fma(m1[4], r4, c15)))));
            const double t2 = m23 * r3;
            const double t3 = fma(m30, r0, fma(m31, r1, m32 * r2));
            const double t4 = m40 * r0;
            r0 = fma(t0, ik, y[0]);
            r1 = fma(t1, ik, y[1]);
            r2 = fma(t2, ik, y[2]);
            r3 = fma(t3, ik, y[3]);
            r4 = fma(t4, ik, y[4]);
        }
        y[0] = r0; y[1] = r1; y[2] = r2; y[3] = r3; y[4] = r4;
    }
#pragma unroll
    for (int r = 0; r < 5; ++r) x[r] = y[r] * pow2(e[r]);
}

// ----------------------------------------------------------------------------------------
// BalancingRiderBicycle gain design: Ackermann's formula == ct.place for a single input
// (dynamics.py:602-615, :1205-1209).  K = e_n^T C^-1 phi(A), evaluated as a ROW VECTOR pushed through the
// factors of phi -- five vector-matrix products instead of matrix-matrix products; A (25 doubles) and a
// few 5-vectors, all in registers.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void br_matrix(const CsfAgentParams& p, double v, double (&A)[5][5]) {
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) A[i][j] = fma(v * v, p.br_A2[i * 5 + j], fma(v, p.br_A1[i * 5 + j], p.br_A0[i * 5 + j]));
}
// w^T A
__device__ __forceinline__ void vec_mat5(const double (&w)[5], const double (&A)[5][5], double (&out)[5]) {
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 5; ++i) t = fma(w[i], A[i][j], t);
        out[j] = t;
    }
}
__device__ __forceinline__ void br_gains(const CsfAgentParams& p, double v, const double* f /* pole features */, double* K) {
    double A[5][5];
    br_matrix(p, v, A);
    // controllability matrix C = [b, Ab, ..., A^4 b]; w^T = e_5^T C^-1  <=>  C^T w = e_5
    double Ct[5][5], col[5], nxt[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) col[i] = p.br_B[i];
#pragma unroll
    for (int c = 0; c < 5; ++c) {
#pragma unroll
        for (int i = 0; i < 5; ++i) Ct[c][i] = col[i];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            double t = 0.0;
#pragma unroll
            for (int j = 0; j < 5; ++j) t = fma(A[i][j], col[j], t);
            nxt[i] = t;
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) col[i] = nxt[i];
    }
    double e[5] = {0, 0, 0, 0, 1}, w[5];
    solve5(Ct, e, w);
    // K = w^T (A - p0 I) (A^2 - 2 re1 A + |p1|^2 I) (A^2 - 2 re2 A + |p2|^2 I)
    double u[5], u2[5];
    vec_mat5(w, A, u);
#pragma unroll
    for (int j = 0; j < 5; ++j) w[j] = fma(-f[0], w[j], u[j]);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const double re = f[1 + 2 * q], im = f[2 + 2 * q];
        vec_mat5(w, A, u);
        vec_mat5(u, A, u2);
        const double m2 = fma(re, re, im * im);
#pragma unroll
        for (int j = 0; j < 5; ++j) w[j] = fma(m2, w[j], fma(-2.0 * re, u[j], u2[j]));
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) K[j] = w[j];
}

// ----------------------------------------------------------------------------------------
// Stochastic rider behaviour: closed-loop poles drawn from the rider-behaviour model, a Gaussian mixture
// over [speed, pole features] conditioned on the speed, in a Yeo-Johnson / log-shifted feature space
// (parameters.py:1398-1402 -> controlbehavior.py:1414-1469 PoleModel.sample_poles, :1337-1412 sample,
// :477-533 _get_conditional_gmm, :962-985 inverse_transform).  Counter-based random numbers
// (Philox-4x32-10; key = seed, counter = road user, draw number): a road user's poles do not depend on how
// the crowd is grouped or sharded, and the host oracle (oracle/pole_sampling.py) reproduces every draw.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t* out) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0;
        c1 = l1;
        c2 = h0 ^ c3 ^ k1;
        c3 = l0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double yj_forward(double x, double lam) {          // sklearn PowerTransformer
    if (x >= 0.0) return fabs(lam) < 2.220446049250313e-16 ? log1p(x) : (pow(x + 1.0, lam) - 1.0) / lam;
    return fabs(lam - 2.0) > 2.220446049250313e-16 ? -(pow(-x + 1.0, 2.0 - lam) - 1.0) / (2.0 - lam) : -log1p(-x);
}
__device__ __forceinline__ double yj_inverse(double y, double lam) {          // NaN outside the transform's range
    if (y >= 0.0) return fabs(lam) < 2.220446049250313e-16 ? exp(y) - 1.0 : pow(y * lam + 1.0, 1.0 / lam) - 1.0;
    return fabs(lam - 2.0) > 2.220446049250313e-16 ? 1.0 - pow(-(2.0 - lam) * y + 1.0, 1.0 / (2.0 - lam)) : 1.0 - exp(-y);
}
// pole features [p0_real, p1_real, p1_imag, p2_real, p2_imag] of road user `agent` at speed v; draws
// *draws, *draws + 1, ... until the sample is inside the range of the inverse transform and stable.
__device__ bool br_sample_poles(const CsfAgentParams& p, double v, unsigned long long agent, int* draws, double* f) {
    const int nc = p.br_n_comp;
    const double xt = (yj_forward(v, p.br_lam[0]) - p.br_sc_mean[0]) / p.br_sc_scale[0];
    double w[4], wsum = 0.0;
    for (int c = 0; c < 4; ++c) {
        w[c] = 0.0;
        if (c < nc) {
            const double d = xt - p.br_mu_g[c];
            w[c] = p.br_w[c] * exp(-0.5 * d * d / p.br_var_g[c]) / sqrt(CSF_TWO_PI * p.br_var_g[c]);
            wsum += w[c];
        }
    }
    bool any0 = false;
    for (int c = 0; c < nc; ++c) { w[c] /= wsum; any0 = any0 || w[c] == 0.0; }
    if (any0) {                                                  // controlbehavior.py:523-526
        double s2 = 0.0;
        for (int c = 0; c < nc; ++c) { if (w[c] == 0.0) w[c] = 2.220446049250313e-16 * nc; s2 += w[c]; }
        for (int c = 0; c < nc; ++c) w[c] /= s2;
    }
    const uint32_t k0 = (uint32_t)(p.br_seed & 0xffffffffull), k1 = (uint32_t)(p.br_seed >> 32);
    const uint32_t a_lo = (uint32_t)(agent & 0xffffffffull), a_hi = (uint32_t)(agent >> 32);
    for (int t = 0; t < 1000; ++t) {
        uint32_t r[8];
        philox4x32_10(a_lo, (uint32_t)(*draws + t), 0u, a_hi, k0, k1, r);
        philox4x32_10(a_lo, (uint32_t)(*draws + t), 1u, a_hi, k0, k1, r + 4);
        double u[8], z[6];
        for (int i = 0; i < 8; ++i) u[i] = ((double)r[i] + 0.5) * 2.3283064365386963e-10;
        for (int i = 0; i < 3; ++i) {
            const double rad = sqrt(-2.0 * log(u[1 + 2 * i]));
            double sn, cs;
            sincos(CSF_TWO_PI * u[2 + 2 * i], &sn, &cs);
            z[2 * i] = rad * cs;
            z[2 * i + 1] = rad * sn;
        }
        int comp = nc - 1;
        double cum = 0.0;
        for (int c = 0; c < nc; ++c) {
            cum += w[c];
            if (u[0] < cum) { comp = c; break; }
        }
        bool ok = true;
        for (int i = 0; i < 5; ++i) {
            double x = p.br_mu[comp][i] + p.br_slope[comp][i] * (xt - p.br_mu_g[comp]);
            for (int j = 0; j <= i; ++j) x += p.br_chol[comp][i * (i + 1) / 2 + j] * z[j];
            double y = yj_inverse(x * p.br_sc_scale[1 + i] + p.br_sc_mean[1 + i], p.br_lam[1 + i]);
            if (p.br_log_sign[i] != 0.0) y = (exp(y) + p.br_log_a[i]) / p.br_log_sign[i];
            f[i] = y;
            ok = ok && isfinite(y);
        }
        ok = ok && f[0] <= 0.0 && f[1] <= 0.0 && f[3] <= 0.0;
        if (ok) { *draws += t + 1; return true; }
    }
    *draws += 1000;
    return false;
}
// poles of a BalancingRider agent for the speed v: the regression of the component mean, or -- stochastic
// behaviour -- the agent's current sample, re-drawn when the speed has moved by more than the threshold
// since the last draw (update_control_params, parameters.py:1376-1411)
__device__ void br_poles_for(const CsfAgentParams& p, const CsfAgentState& st, int64_t k, double v, double* f, int* flags) {
    if (!p.br_stochastic) {
        for (int i = 0; i < 5; ++i) f[i] = p.br_pole_icpt[i] + p.br_pole_coef[i] * v;
        return;
    }
    if (fabs(v - st.br_vlast[k]) > p.br_resample_thresh) {
        int draws = st.br_draws[k];
        if (!br_sample_poles(p, v, (unsigned long long)st.br_stream[k], &draws, f)) *flags |= 16;
        st.br_draws[k] = draws;
        st.br_vlast[k] = v;
        for (int i = 0; i < 5; ++i) st.br_poles[(size_t)i * st.n + k] = f[i];
    } else {
        for (int i = 0; i < 5; ++i) f[i] = st.br_poles[(size_t)i * st.n + k];
    }
}
// gains of every agent for its current speed (BalancingRiderDynamics.__init__ -> _get_gains(v),
// dynamics.py:305-306, :602-615): also draws the first poles of a stochastic rider
__global__ void br_init_kernel(CsfAgentState st, CsfAgentParams p) {
    const int64_t k = st.first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= st.first + st.count) return;
    double f[5], g[5];
    int flags = 0;
    const double v = st.dyn_v[k];
    br_poles_for(p, st, k, v, f, &flags);
    br_gains(p, v, f, g);
    for (int r = 0; r < 5; ++r) st.br_gains[(size_t)r * st.n + k] = g[r];
    if (flags && st.status != nullptr) {
        atomicOr(st.status, flags);
        if (st.status_host != nullptr) *reinterpret_cast<volatile int32_t*>(st.status_host) = 1;
    }
}

// ----------------------------------------------------------------------------------------
// the per-agent kernel
// ----------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T ldT(const void* p, int64_t k) { return reinterpret_cast<const T*>(p)[k]; }
template <typename T> __device__ __forceinline__ void stT(void* p, int64_t k, T v) { reinterpret_cast<T*>(p)[k] = v; }

// Everything one agent does in a step; returns true if a new payload element was produced (*out).
// `pair_done()` is called exactly once, after everything that does not depend on the pair kernel of this
// step (state and queue loads, navigation machine, destination force) and before the repulsive force is
// read: the fused step launches this kernel as a programmatic dependent of the pair kernel, so that part
// runs in the pair kernel's tail, on the SMs that have already run out of items.
template <typename T, int MODEL, int MODE, typename PairDone>
__device__ __forceinline__ bool agent_body(const CsfAgentState& st, const CsfAgentParams& p, int64_t n_total,
                                           const T* __restrict__ frep, const T* __restrict__ froad,
                                           T* __restrict__ force, T* __restrict__ fdest_out,
                                           void* __restrict__ next_xycs, const CsfStepFusion& fu, int64_t k,
                                           Xycs<T>* out, PairDone pair_done) {
    Agent<T> a;
    a.x = st.x[k];
    a.y = st.y[k];
    a.psi = ldT<T>(st.psi, k);
    a.v = ldT<T>(st.v, k);
    a.delta = (MODEL != CSF_MODEL_PLANARPOINT) ? ldT<T>(st.delta, k) : (T)0;
    a.theta = (MODEL == CSF_MODEL_INVPENDULUM || MODEL == CSF_MODEL_BALANCINGRIDER) ? ldT<T>(st.theta, k) : (T)0;
    a.vd_def = ldT<T>(st.vd_default, k);
    a.i = st.step_i[k];
    a.ptr = st.dest_ptr[k];
    a.qlen = st.dest_len[k];
    a.znav = st.znav[k];
    a.z_v0 = ldT<T>(st.znav_v0, k);
    a.z_d0 = ldT<T>(st.znav_d0, k);
    a.z_d1 = ldT<T>(st.znav_d1, k);
    a.q = st.destq + (size_t)k * p.q_cap * 3;
    a.flags = 0;
    bool produced = false;
    // everything else the step reads from global memory, issued before the first use (independent loads:
    // one round trip instead of a chain of them)
    constexpr bool kHist = MODEL == CSF_MODEL_TWOD || MODEL == CSF_MODEL_INVPENDULUM || MODEL == CSF_MODEL_PLANARPOINT;
    double pv_x = 0.0, pv_y = 0.0;
    int hist_step = 0;
    if (kHist) {
        pv_x = st.prev_x[k];
        pv_y = st.prev_y[k];
        hist_step = st.hist_step[k];
    }
    // dynamic state of the richer models (K3 needs it ~2,000 instructions from here: fetched now)
    int pre_zr = 0, pre_run = 0;
    double pre_x[5] = {0, 0, 0, 0, 0}, pre_g[5] = {0, 0, 0, 0, 0}, pre_v = 0.0;
    if (MODE != MODE_FORCES) {
        if (MODEL == CSF_MODEL_INVPENDULUM) {
            pre_zr = st.ip_zrid[k];
            pre_run = st.ip_delta_run[k];
#pragma unroll
            for (int r = 0; r < 5; ++r) pre_x[r] = st.ip_x[(size_t)r * st.n + k];
        } else if (MODEL == CSF_MODEL_BALANCINGRIDER) {
            pre_v = st.dyn_v[k];
#pragma unroll
            for (int r = 0; r < 5; ++r) {
                pre_x[r] = st.dyn_x[(size_t)r * st.n + k];
                pre_g[r] = st.br_gains[(size_t)r * st.n + k];
            }
        } else if (MODEL == CSF_MODEL_PLANARPOINT) {
            pre_v = st.dyn_v[k];
            pre_x[0] = st.dyn_x[k];
        }
    }
    T frx0 = (T)0, fry0 = (T)0, fox = (T)0, foy = (T)0;
    const bool fused_rep = MODE == MODE_STEP && fu.partial != nullptr;
    const bool have_rep = MODE != MODE_ADVANCE && n_total > 1 && (frep != nullptr || fused_rep);
    load_window(a);

    T Fx, Fy;
    if (MODE != MODE_ADVANCE) {
        // ---- K2: destination force + assembly (intersection.py:797-799, :841-862) ----
        T fdx, fdy;
        destination_force<T, MODEL>(a, p, st, k, pv_x, pv_y, fdx, fdy);
        pair_done();
        if (froad != nullptr) {       // (written by the road kernels of this step: read behind the dependency wait too)
            fox = froad[k * 2];
            foy = froad[k * 2 + 1];
        }
        if (have_rep && fused_rep) {
            // partial sums of the tiled pair kernel, one slab per chunk group, reduced here in fixed order
            // (what reduce_groups_kernel does as a launch of its own)
            const T* part = reinterpret_cast<const T*>(fu.partial) + (size_t)(fu.partial_offset + k) * 2;
            for (int g = 0; g < fu.n_groups; ++g) {
                frx0 += part[(size_t)g * fu.partial_stride * 2];
                fry0 += part[(size_t)g * fu.partial_stride * 2 + 1];
            }
            frx0 *= (T)fu.f0;
            fry0 *= (T)fu.f0;
        } else if (have_rep) {
            frx0 = frep[k * 2];
            fry0 = frep[k * 2 + 1];
        }
        T frx = (T)0, fry = (T)0;
        if (have_rep) {
            frx = frx0;
            fry = fry0;
            const T rin = sqrt(frx * frx + fry * fry), r = sqrt(fdx * fdx + fdy * fdy);
            if (rin > r) {  // utils.limitMagnitude, utils.py:79-84
                frx = frx * r / rin;
                fry = fry * r / rin;
            }
        }
        Fx = frx + fdx;
        Fy = fry + fdy;
        Fx += fox;
        Fy += foy;
        if (force != nullptr) { force[k * 2] = Fx; force[k * 2 + 1] = Fy; }
        if (fdest_out != nullptr) { fdest_out[k * 2] = fdx; fdest_out[k * 2 + 1] = fdy; }
        st.dest_ptr[k] = a.ptr;
        st.znav[k] = a.znav;
        stT<T>(st.znav_v0, k, a.z_v0);
        stT<T>(st.znav_d0, k, a.z_d0);
        stT<T>(st.znav_d1, k, a.z_d1);
        if (!(isfinite((double)Fx) && isfinite((double)Fy))) a.flags |= 1;
    } else {
        pair_done();
        Fx = force[k * 2];
        Fy = force[k * 2 + 1];
    }

    if (MODE != MODE_FORCES) {
        // ---- K3: rider control + dynamics ----
        const double ox = a.x, oy = a.y;
        bool wrap = true;
        if (MODEL == CSF_MODEL_TWOD || MODEL == CSF_MODEL_BICYCLE) {
            if (MODEL == CSF_MODEL_TWOD && (a.znav & 4)) {  // vehicle.py:1397-1398
                a.v = (T)0;
                a.delta = (T)0;
            } else control_move(a, p, Fx, Fy);
        } else if (MODEL == CSF_MODEL_INVPENDULUM) {
            // updateRidingState, vehicle.py:1932-1950
            int zr = pre_zr;
            const int run = pre_run;
            const bool cvwalk = a.v < (T)p.v_max_walk;
            const bool cdelta = run >= min(a.i, p.hist_len) + 1;
            const bool ride = !cvwalk && (((zr & 2) && cdelta) || (zr & 1));
            zr = ride ? 1 : 2;
            st.ip_zrid[k] = zr;
            double xs[5];
#pragma unroll
            for (int r = 0; r < 5; ++r) xs[r] = pre_x[r];
            if (a.znav & 4) {  // :1898-1899
                a.v = (T)0; a.delta = (T)0; a.theta = (T)0;
            } else if (ride) {
                // step_pos (:1850-1881): P-controlled speed, Euler position with the OLD psi
                const T vdF = sqrt(Fx * Fx + Fy * Fy);
                const T acc = clampT((T)p.k_p_v * (vdF - a.v), (T)p.a_max[0], (T)p.a_max[1]);
                const T v = clampT(a.v + (T)p.t_s * acc, (T)p.v_max_riding[0], (T)p.v_max_riding[1]);
                T sn, cs;
                sincosT(a.psi, &sn, &cs);
                a.y += (double)((T)p.t_s * v * sn);
                a.x += (double)((T)p.t_s * v * cs);
                a.v = v;
                // step_yaw (:1810-1848) with the updated speed
                invpend_yaw_step(p, (double)a.v, (double)atan2(Fy, Fx), xs);
                a.psi = (T)limit_angle(xs[4]);
                a.delta = (T)limit_angle(xs[0]);
                a.theta = (T)limit_angle(xs[2]);
            } else {  // walking, :1905-1916
                a.v = (T)p.v_max_walk;
                a.theta = (T)0;
                control_move(a, p, Fx, Fy);
                xs[0] = (double)a.delta; xs[1] = 0.0; xs[2] = (double)a.theta; xs[3] = 0.0; xs[4] = (double)a.psi;
            }
            for (int r = 0; r < 5; ++r) st.ip_x[(size_t)r * st.n + k] = xs[r];
            const bool ok = fabs(a.delta) < (T)p.delta_max_walk;
            // the window restarts when Vehicle.i wraps to 0 (traj[4, 0:i+1])
            const int inew = (a.i + 1) % p.traj_len;
            st.ip_delta_run[k] = ok ? ((inew == 0) ? 1 : run + 1) : 0;
            stT<T>(st.theta, k, a.theta);
        } else if (MODEL == CSF_MODEL_PLANARPOINT) {
            // PlanarPointDynamics.step, dynamics.py:1051-1079 (closed-form implicit midpoint)
            wrap = false;
            const double vold = pre_v;
            const double vdF = sqrt((double)Fx * Fx + (double)Fy * Fy);
            const double acc = clampT(p.k_p_v * (vdF - vold), p.a_max[0], p.a_max[1]);
            const double v = clampT(vold + p.t_s * acc, p.v_max_riding[0], p.v_max_riding[1]);
            const double vbar = 0.5 * (v + vold);  // vehicle.s[3] == dynamics.v in the reference
            const double psi_c = limit_angle(atan2((double)Fy, (double)Fx));  // dynamics.py:112-121
            const double h = p.t_s, kp = p.k_psi;
            const double ps = pre_x[0];
            const double pn = ((1.0 - h * kp / 2) * ps + h * kp * psi_c) / (1.0 + h * kp / 2);
            double sn, cs;
            sincos(0.5 * (ps + pn), &sn, &cs);
            a.x += h * vbar * cs;
            a.y += h * vbar * sn;
            st.dyn_x[k] = pn;
            st.dyn_v[k] = v;
            a.psi = (T)limit_angle(pn);
            a.v = (T)v;
        } else if (MODEL == CSF_MODEL_BALANCINGRIDER) {
            // BalancingRiderDynamics.step, dynamics.py:674-705 (closed-form implicit midpoint)
            wrap = false;
            const double vold = pre_v;
            const double vdF = sqrt((double)Fx * Fx + (double)Fy * Fy);
            const double acc = clampT(p.k_p_v * (vdF - vold), p.a_max[0], p.a_max[1]);
            const double v = clampT(vold + p.t_s * acc, p.v_max_riding[0], p.v_max_riding[1]);
            const double vbar = 0.5 * (v + vold);  // vehicle.s[3] == dynamics.v in the reference
            double g[5], xb[5];
#pragma unroll
            for (int r = 0; r < 5; ++r) xb[r] = pre_x[r];
            if (v != vold && !p.br_fixed_gains) {  // :680-681 (fixed gains: dynamics.py:606-607)
                double pf[5];
                br_poles_for(p, st, k, vbar, pf, &a.flags);
                br_gains(p, vbar, pf, g);
                for (int r = 0; r < 5; ++r) st.br_gains[(size_t)r * st.n + k] = g[r];
            } else
#pragma unroll
                for (int r = 0; r < 5; ++r) g[r] = pre_g[r];
            const double psi_F = limit_angle(atan2(-(double)Fy, (double)Fx));  // :661-671
            const double psi_c = xb[4] + angle_difference(xb[4], psi_F);
            double A[5][5], L[5][5], rhs[5], xn[5];
            br_matrix(p, vbar, A);
            const double h = p.t_s;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                double s = h * p.br_B[i] * g[4] * psi_c;
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const double ac = A[i][j] - p.br_B[i] * g[j];          // A_c
                    L[i][j] = ((i == j) ? 1.0 : 0.0) - 0.5 * h * ac;
                    s += (((i == j) ? 1.0 : 0.0) + 0.5 * h * ac) * xb[j];
                }
                rhs[i] = s;
            }
            solve5(L, rhs, xn);
            double sn, cs;
            sincos(0.5 * (xb[4] + xn[4]), &sn, &cs);
            // bike frame (N): p_x = x, p_y = -y   (dynamics.py:347-356, :389-397)
            a.x += h * vbar * cs;
            a.y -= h * vbar * sn;
            for (int r = 0; r < 5; ++r) st.dyn_x[(size_t)r * st.n + k] = xn[r];
            st.dyn_v[k] = v;
            a.psi = (T)(-limit_angle(xn[4]));
            a.v = (T)v;
            a.delta = (T)(-limit_angle(xn[1]));
            a.theta = (T)limit_angle(xn[0]);
            stT<T>(st.theta, k, a.theta);
            stT<T>(st.deltadot, k, (T)(-xn[3]));
            stT<T>(st.thetadot, k, (T)xn[2]);
        }
        // ---- bookkeeping: Vehicle.i, traj ring (vehicle.py:1407-1410, :320-321) ----
        a.i = a.i + 1;
        if (wrap) a.i %= p.traj_len;
        st.step_i[k] = a.i;
        st.x[k] = a.x;
        st.y[k] = a.y;
        stT<T>(st.psi, k, a.psi);
        stT<T>(st.v, k, a.v);
        if (MODEL != CSF_MODEL_PLANARPOINT) stT<T>(st.delta, k, a.delta);
        if (MODEL != CSF_MODEL_BALANCINGRIDER && MODEL != CSF_MODEL_BICYCLE) {
            st.prev_x[k] = ox;
            st.prev_y[k] = oy;
            const int hs = hist_step + 1;
            st.hist_step[k] = hs;
            const int row = hs & (p.hist_cap - 1);
            st.hist_x[(size_t)row * st.n + k] = a.x;
            st.hist_y[(size_t)row * st.n + k] = a.y;
        }
        if (!(isfinite(a.x) && isfinite(a.y) && isfinite((double)a.psi) && isfinite((double)a.v))) a.flags |= 1;
        if (next_xycs != nullptr) {
            int ovf = 0;
            *out = store_xycs<T>(next_xycs, st.payload_offset + k, a.x, a.y, a.psi, 1.0 / p.q_scale, p.q_origin[0],
                                 p.q_origin[1], &ovf);
            if (ovf) a.flags |= 4;
            produced = true;
        }
    }
    if (a.flags && st.status != nullptr) {
        atomicOr(st.status, a.flags);
        // host-mapped mirror: the host polls it without a copy or a synchronisation
        if (st.status_host != nullptr) *reinterpret_cast<volatile int32_t*>(st.status_host) = 1;
    }
    return produced;
}

// One thread per agent.  With `fu.comm` (a crowd sharded over the GPUs of a node, csf_peer.cu) the kernel
// is also the step's side of the payload exchange: block 0 tells every peer that this rank has finished
// reading their payload entries (the pair kernel ran before this one); after the agents have been
// stepped every block waits for the peers' same signal, stores its agents' new payload elements
// straight into every peer's buffer (16-byte NVLink stores), and the last block to finish raises the
// data flags.  No separate wait / signal / push launches, no collective call.
template <typename T, int MODEL, int MODE>
__global__ void __launch_bounds__(128) agent_kernel(CsfAgentState st, CsfAgentParams p, int64_t n_total,
                                                    const T* __restrict__ frep, const T* __restrict__ froad,
                                                    T* __restrict__ force, T* __restrict__ fdest_out,
                                                    void* __restrict__ next_xycs, CsfStepFusion fu) {
    const int64_t k = st.first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = k < st.first + st.count;
    const bool peers = MODE == MODE_STEP && fu.comm.world > 1;
    __shared__ uint32_t s_epoch, s_last;
    if (peers) {
        if (threadIdx.x == 0) s_epoch = ld_sys(fu.comm.seq + SEQ_PUSH) + 1u;
        __syncthreads();
    }
    // The pair kernel of this step has completed (all of its memory operations are visible): a no-op unless
    // the kernel was launched as a programmatic dependent.  Only then may the peers be told that this rank
    // has finished reading their payload entries.
    auto pair_done = [&]() {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (peers && blockIdx.x == 0 && threadIdx.x < fu.comm.world && (int)threadIdx.x != fu.comm.rank)
            st_sys(fu.comm.read_flags[threadIdx.x] + fu.comm.rank, s_epoch);
    };
    Xycs<T> mine;
    bool produced = false;
    if (active) produced = agent_body<T, MODEL, MODE>(st, p, n_total, frep, froad, force, fdest_out, next_xycs, fu, k, &mine, pair_done);
    else pair_done();
    if (peers) {
        const CsfPeerComm& c = fu.comm;
        if (threadIdx.x < c.world && (int)threadIdx.x != c.rank && peer_ok(c)) {
            if (!spin_ge(c.read_flags[c.rank] + threadIdx.x, s_epoch)) peer_fail(c, 2u);
        }
        __syncthreads();
        if (peer_ok(c)) {
            if (produced) {
                const int64_t idx = st.payload_offset + k;
                for (int q = 0; q < c.world; ++q)
                    if (q != c.rank) reinterpret_cast<Xycs<T>*>(c.payload[q])[idx] = mine;
            }
            __threadfence_system();
            __syncthreads();
            if (threadIdx.x == 0) {
                const uint32_t t = atomicAdd(c.seq + SEQ_BLOCKS, 1u);
                s_last = (t == gridDim.x - 1) ? 1u : 0u;
                if (s_last) {
                    c.seq[SEQ_BLOCKS] = 0;
                    c.seq[SEQ_PUSH] = s_epoch;
                    __threadfence_system();
                }
            }
            __syncthreads();
            if (s_last && threadIdx.x < c.world && (int)threadIdx.x != c.rank)
                st_sys(c.data_flags[threadIdx.x] + c.rank, s_epoch);
        }
    }
}

template <typename T>
__global__ void pack_state_kernel(CsfAgentState st, double inv_q, double ox, double oy, void* xycs) {
    const int64_t k = st.first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= st.first + st.count) return;
    int ovf = 0;
    store_xycs<T>(xycs, st.payload_offset + k, st.x[k], st.y[k], ldT<T>(st.psi, k), inv_q, ox, oy, &ovf);
    if (ovf && st.status != nullptr) {
        atomicOr(st.status, 4);
        if (st.status_host != nullptr) *reinterpret_cast<volatile int32_t*>(st.status_host) = 1;
    }
}
template <typename T>
__global__ void pack_xypsi_kernel(const double* x, const double* y, const double* psi, int64_t n, double inv_q,
                                  double ox, double oy, void* xycs) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int ovf = 0;
    store_xycs<T>(xycs, k, x[k], y[k], (T)psi[k], inv_q, ox, oy, &ovf);
}

template <typename T, int MODE>
int launch_agent(int model, const CsfAgentState* st, const CsfAgentParams* p, int64_t n_total, const T* frep,
                 const T* froad, T* force, T* fdest, void* next_xycs, cudaStream_t stream,
                 const CsfStepFusion* fusion = nullptr) {
    CsfStepFusion fu;
    memset(&fu, 0, sizeof(fu));
    if (fusion) fu = *fusion;
    if (st->count <= 0 && fu.comm.world <= 1) return 0;
    const unsigned grid = (unsigned)((st->count + 127) / 128) > 0 ? (unsigned)((st->count + 127) / 128) : 1u;
    // fu.pdl: programmatic dependent launch -- the kernel may start while the preceding kernel in the stream (the
    // pair kernel, which releases its dependents when its CTAs run out of items) is still draining; everything
    // that depends on it sits behind griddepcontrol.wait (pair_done above)
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = fu.pdl ? 1 : 0;
#define CSF_LAUNCH(MODEL)                                                                                      \
    cudaLaunchKernelEx(&cfg, agent_kernel<T, MODEL, MODE>, *st, *p, n_total, frep, froad, force, fdest, next_xycs, fu)
    switch (model) {
        case CSF_MODEL_TWOD: CSF_LAUNCH(CSF_MODEL_TWOD); break;
        case CSF_MODEL_INVPENDULUM: CSF_LAUNCH(CSF_MODEL_INVPENDULUM); break;
        case CSF_MODEL_BALANCINGRIDER: CSF_LAUNCH(CSF_MODEL_BALANCINGRIDER); break;
        case CSF_MODEL_PLANARPOINT: CSF_LAUNCH(CSF_MODEL_PLANARPOINT); break;
        case CSF_MODEL_BICYCLE: CSF_LAUNCH(CSF_MODEL_BICYCLE); break;
        default:
            csf_set_error("csf_agent_*: unknown model id", cudaErrorInvalidValue);
            return -(int)cudaErrorInvalidValue;
    }
#undef CSF_LAUNCH
    CSF_CHECK_LAUNCH("agent_kernel");
    return 0;
}

}  // namespace

extern "C" {

int csf_agent_forces_f32(int model, const CsfAgentState* st, const CsfAgentParams* p, int64_t n_total,
                         const float* frep, const float* froad, float* force, float* fdest, csf_stream_t s) {
    return launch_agent<float, MODE_FORCES>(model, st, p, n_total, frep, froad, force, fdest, nullptr, (cudaStream_t)s);
}
int csf_agent_forces_f64(int model, const CsfAgentState* st, const CsfAgentParams* p, int64_t n_total,
                         const double* frep, const double* froad, double* force, double* fdest, csf_stream_t s) {
    return launch_agent<double, MODE_FORCES>(model, st, p, n_total, frep, froad, force, fdest, nullptr, (cudaStream_t)s);
}
int csf_agent_advance_f32(int model, const CsfAgentState* st, const CsfAgentParams* p, const float* force,
                          void* next_xycs, csf_stream_t s) {
    return launch_agent<float, MODE_ADVANCE>(model, st, p, 0, nullptr, nullptr, const_cast<float*>(force), nullptr,
                                             next_xycs, (cudaStream_t)s);
}
int csf_agent_advance_f64(int model, const CsfAgentState* st, const CsfAgentParams* p, const double* force,
                          void* next_xycs, csf_stream_t s) {
    return launch_agent<double, MODE_ADVANCE>(model, st, p, 0, nullptr, nullptr, const_cast<double*>(force), nullptr,
                                              next_xycs, (cudaStream_t)s);
}
int csf_agent_step_f32(int model, const CsfAgentState* st, const CsfAgentParams* p, int64_t n_total,
                       const float* frep, const float* froad, float* force, void* next_xycs, csf_stream_t s) {
    return launch_agent<float, MODE_STEP>(model, st, p, n_total, frep, froad, force, nullptr, next_xycs, (cudaStream_t)s);
}
int csf_agent_step_f64(int model, const CsfAgentState* st, const CsfAgentParams* p, int64_t n_total,
                       const double* frep, const double* froad, double* force, void* next_xycs, csf_stream_t s) {
    return launch_agent<double, MODE_STEP>(model, st, p, n_total, frep, froad, force, nullptr, next_xycs, (cudaStream_t)s);
}
int csf_agent_step_fused_f32(int model, const CsfAgentState* st, const CsfAgentParams* p, int64_t n_total,
                             const CsfStepFusion* fusion, const float* froad, float* force, void* next_xycs,
                             csf_stream_t s) {
    return launch_agent<float, MODE_STEP>(model, st, p, n_total, nullptr, froad, force, nullptr, next_xycs,
                                          (cudaStream_t)s, fusion);
}
int csf_agent_step_fused_f64(int model, const CsfAgentState* st, const CsfAgentParams* p, int64_t n_total,
                             const CsfStepFusion* fusion, const double* froad, double* force, void* next_xycs,
                             csf_stream_t s) {
    return launch_agent<double, MODE_STEP>(model, st, p, n_total, nullptr, froad, force, nullptr, next_xycs,
                                           (cudaStream_t)s, fusion);
}
int csf_br_init(const CsfAgentState* st, const CsfAgentParams* p, csf_stream_t s) {
    if (st->count <= 0) return 0;
    br_init_kernel<<<(unsigned)((st->count + 63) / 64), 64, 0, (cudaStream_t)s>>>(*st, *p);
    CSF_CHECK_LAUNCH("br_init_kernel");
    return 0;
}
int csf_pack_xycs_f32(const CsfAgentState* st, const CsfAgentParams* p, void* xycs, csf_stream_t s) {
    if (st->count <= 0) return 0;
    pack_state_kernel<float><<<(unsigned)((st->count + 127) / 128), 128, 0, (cudaStream_t)s>>>(*st, 1.0 / p->q_scale, p->q_origin[0], p->q_origin[1], xycs);
    CSF_CHECK_LAUNCH("pack_state_kernel");
    return 0;
}
int csf_pack_xycs_f64(const CsfAgentState* st, const CsfAgentParams* p, void* xycs, csf_stream_t s) {
    if (st->count <= 0) return 0;
    pack_state_kernel<double><<<(unsigned)((st->count + 127) / 128), 128, 0, (cudaStream_t)s>>>(*st, 1.0, 0.0, 0.0, xycs);
    CSF_CHECK_LAUNCH("pack_state_kernel");
    return 0;
}
int csf_pack_xypsi_f32(const double* x, const double* y, const double* psi, int64_t n, double q_scale,
                       double origin_x, double origin_y, void* xycs, csf_stream_t s) {
    if (n <= 0) return 0;
    pack_xypsi_kernel<float><<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)s>>>(x, y, psi, n, 1.0 / q_scale, origin_x, origin_y, xycs);
    CSF_CHECK_LAUNCH("pack_xypsi_kernel");
    return 0;
}
int csf_pack_xypsi_f64(const double* x, const double* y, const double* psi, int64_t n, double q_scale,
                       double origin_x, double origin_y, void* xycs, csf_stream_t s) {
    (void)q_scale; (void)origin_x; (void)origin_y;
    if (n <= 0) return 0;
    pack_xypsi_kernel<double><<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)s>>>(x, y, psi, n, 1.0, 0.0, 0.0, xycs);
    CSF_CHECK_LAUNCH("pack_xypsi_kernel");
    return 0;
}

}  // extern "C"
