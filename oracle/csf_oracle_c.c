/*
 * csf_oracle_c.c -- plain C (OpenMP) restatement of the TwoDBicycle stepping path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the CPU baseline that bench.py
 * times on the host cores ("port"), validated against oracle/csf_oracle.py (which is
 * pinned to the reference) in tests/test_oracle_c.py.  Never linked into the product.
 *
 * Follows, like the numpy oracle, the reference's own formulation (trigonometric
 * pair force, angleDifference-based field-of-view mask), float64 throughout:
 *   pair force   vehicle.py:1584-1648, mask intersection.py:706-741
 *   limitAngle / angleDifference / cart2polar   utils.py:124-194
 *   updateDestination / updateNavState          vehicle.py:545-594, :354-457
 *   TwoDBicycle.calcDestinationForce            vehicle.py:1443-1558
 *     (splprep/splev = FITPACK interpolating cubic B-spline, restated)
 *   clip + sum                                  intersection.py:841-848, utils.py:56-86
 *   Bicycle.control / move, TwoDBicycle.step    vehicle.py:1218-1272, :1386-1414
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.14159265358979323846
#define TWO_PI (2.0 * PI)

static double limit_angle(double th) {
    th = floor(th / TWO_PI) * (-TWO_PI) + th;
    if (th > PI) th -= TWO_PI;
    else if (th < -PI) th += TWO_PI;
    return th;
}
static double angle_difference(double a1, double a2) {
    double da = a1 > a2 ? a1 - a2 : a2 - a1;
    if (da > PI) da = TWO_PI - da;
    double t1 = fabs(limit_angle(a1 - da) - a2), t2 = fabs(limit_angle(a1 + da) - a2);
    return t1 < t2 ? -da : da;
}
static double clampd(double x, double lo, double hi) { return fmax(fmin(x, hi), lo); }
static double sgn(double x) { return (x > 0) - (x < 0); }

/* torchrun exports OMP_NUM_THREADS=1 to its workers and the OpenMP runtime of the process reads it once:
 * the CPU baseline sets the thread count explicitly */
void csf_c_set_num_threads(int n) {
    if (n > 0) omp_set_num_threads(n);
}
int csf_c_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* fp = [f_0, e_0, e_1, sigma_0..3, hfov].  out[t] = sum over sources i of mask(i, j_t) F(i -> j_t). */
void csf_c_pair_forces(int64_t n, const double* x, const double* y, const double* psi, const double* fp, int p2r,
                       int64_t nt, const int64_t* tgt, double* out) {
    const double f0 = fp[0], e0 = fp[1], e1 = fp[2], s0 = fp[3], s1 = fp[4], s2_ = fp[5], s3 = fp[6];
    const double hh = fp[7] / 2;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t t = 0; t < nt; ++t) {
        const int64_t j = tgt[t];
        double ax = 0.0, ay = 0.0;
        for (int64_t i = 0; i < n; ++i) {
            if (i == j || f0 == 0.0) continue;
            /* mask: source i in the field of view of target j (hfov of the source's class) */
            const double az = limit_angle(atan2(y[i] - y[j], x[i] - x[j]));
            const double rel = angle_difference(psi[j], az);
            if (fabs(rel) > hh) continue;
            if (p2r && rel > 0) continue;
            /* field of source i at target j */
            const double psi_rel = psi[i] - psi[j];
            const double sn = sin(psi_rel), sin2 = sn * sn;
            const double vd0 = s0 + s1 * sin2, vd1 = s2_ + s3 * sin2, e = e0 - e1 * sin2;
            const double dx = x[j] - x[i], dy = y[j] - y[i];
            const double rho = sqrt(dx * dx + dy * dy);
            double phi1 = acos(dx / rho);
            if (dy < 0) phi1 = -phi1;
            const double phi = limit_angle(phi1 - psi[i]);
            const double cp = cos(phi), sp = sin(phi);
            const double sigma = vd0 - vd1 * sqrt((1 - cp) / 2);
            const double dsigm = -vd1 * sqrt((1 + cp) / 2) * sgn(phi) / 2;
            const double q = sqrt(1 - (e * cp) * (e * cp));
            const double P = f0 * exp(-rho * q / sigma);
            const double Frho = q / sigma;
            const double Fphi = -((1 - (e * cp) * (e * cp)) * dsigm - e * e * sp * cp * sigma) / (sigma * sigma * q);
            const double c1 = cos(phi1), s1_ = sin(phi1);
            const double Fx = Frho * c1 - Fphi * s1_, Fy = Frho * s1_ + Fphi * c1;
            const double F = sqrt(Fx * Fx + Fy * Fy);
            ax += P * Fx / F;
            ay += P * Fy / F;
        }
        out[t * 2] = ax;
        out[t * 2 + 1] = ay;
    }
}

/* ---- interpolating cubic B-spline through m in {4,5,6} points (FITPACK splprep s=0) ---------- */
typedef struct { int m; double t[10]; double cx[6], cy[6]; } spline_t;

static void basis(const spline_t* s, double u, int l, double* N3, double* N2, double* N1) {
    double left[4], right[4], N[4];
    for (int j = 1; j <= 3; ++j) { left[j] = u - s->t[l + 1 - j]; right[j] = s->t[l + j] - u; }
    N[0] = 1.0;
    for (int j = 1; j <= 3; ++j) {
        double saved = 0.0;
        for (int r = 0; r < j; ++r) {
            const double temp = N[r] / (right[r + 1] + left[j - r]);
            N[r] = saved + right[r + 1] * temp;
            saved = left[j - r] * temp;
        }
        N[j] = saved;
        if (j == 1) { N1[0] = N[0]; N1[1] = N[1]; }
        if (j == 2) { N2[0] = N[0]; N2[1] = N[1]; N2[2] = N[2]; }
    }
    for (int j = 0; j < 4; ++j) N3[j] = N[j];
}
static int spline_fit(spline_t* s, const double* px, const double* py, int m) {
    double u[6];
    s->m = m;
    u[0] = 0.0;
    for (int k = 1; k < m; ++k) {
        const double d = hypot(px[k] - px[k - 1], py[k] - py[k - 1]);
        if (!(d > 0.0)) return 0;
        u[k] = u[k - 1] + d;
    }
    for (int k = 1; k < m; ++k) u[k] /= u[m - 1];
    u[m - 1] = 1.0;
    for (int k = 0; k < 4; ++k) { s->t[k] = 0.0; s->t[m + k] = 1.0; }
    for (int k = 0; k < m - 4; ++k) s->t[4 + k] = u[2 + k];
    double A[6][6], bx[6], by[6];
    memset(A, 0, sizeof(A));
    for (int i = 0; i < m; ++i) { bx[i] = px[i]; by[i] = py[i]; }
    A[0][0] = 1.0;
    A[m - 1][m - 1] = 1.0;
    for (int i = 1; i <= m - 2; ++i) {
        int l = i + 2 < m - 1 ? i + 2 : m - 1;
        double N3[4], N2[3], N1[2];
        basis(s, u[i], l, N3, N2, N1);
        for (int r = 0; r < 4; ++r) A[i][l - 3 + r] = N3[r];
    }
    for (int k = 0; k < m; ++k) {          /* partial pivoting */
        int piv = k;
        for (int i = k + 1; i < m; ++i) if (fabs(A[i][k]) > fabs(A[piv][k])) piv = i;
        if (piv != k) {
            for (int j = 0; j < m; ++j) { double t = A[k][j]; A[k][j] = A[piv][j]; A[piv][j] = t; }
            double t = bx[k]; bx[k] = bx[piv]; bx[piv] = t;
            t = by[k]; by[k] = by[piv]; by[piv] = t;
        }
        for (int i = k + 1; i < m; ++i) {
            const double f = A[i][k] / A[k][k];
            for (int j = k; j < m; ++j) A[i][j] -= f * A[k][j];
            bx[i] -= f * bx[k];
            by[i] -= f * by[k];
        }
    }
    for (int k = m - 1; k >= 0; --k) {
        double sx = bx[k], sy = by[k];
        for (int j = k + 1; j < m; ++j) { sx -= A[k][j] * s->cx[j]; sy -= A[k][j] * s->cy[j]; }
        s->cx[k] = sx / A[k][k];
        s->cy[k] = sy / A[k][k];
    }
    return 1;
}
static void spline_eval(const spline_t* s, double u, double* o /* x y dx dy ddx ddy */) {
    int l = 3;
    while (u >= s->t[l + 1] && l != s->m - 1) ++l;
    double N3[4], N2[3], N1[2], d1x[3], d1y[3];
    basis(s, u, l, N3, N2, N1);
    const double* cx = s->cx + (l - 3);
    const double* cy = s->cy + (l - 3);
    o[0] = o[1] = 0.0;
    for (int r = 0; r < 4; ++r) { o[0] += N3[r] * cx[r]; o[1] += N3[r] * cy[r]; }
    for (int r = 0; r < 3; ++r) {
        const double w = 3.0 / (s->t[l + 1 + r] - s->t[l - 2 + r]);
        d1x[r] = w * (cx[r + 1] - cx[r]);
        d1y[r] = w * (cy[r + 1] - cy[r]);
    }
    o[2] = N2[0] * d1x[0] + N2[1] * d1x[1] + N2[2] * d1x[2];
    o[3] = N2[0] * d1y[0] + N2[1] * d1y[1] + N2[2] * d1y[2];
    const double w0 = 2.0 / (s->t[l + 1] - s->t[l - 1]), w1 = 2.0 / (s->t[l + 2] - s->t[l]);
    o[4] = N1[0] * w0 * (d1x[1] - d1x[0]) + N1[1] * w1 * (d1x[2] - d1x[1]);
    o[5] = N1[0] * w0 * (d1y[1] - d1y[0]) + N1[1] * w1 * (d1y[2] - d1y[1]);
}

/* ---- per-agent state of a TwoDBicycle sub-group (struct of arrays, length ns) ------------------- */
typedef struct {
    int64_t ns, qcap;
    double *s;          /* [ns][5]  x y psi v delta */
    int32_t *i, *ptr, *znav;
    const int32_t *qlen;
    const double *destq; /* [ns][qcap][3] */
    double *znavp;      /* [ns][3] v0 d0 d1 */
    const double *vd;   /* [ns] */
    double *prev;       /* [ns][2] */
    double *hist;       /* [ns][128][2] */
    int32_t *hstep;
} twod_t;

/* p = [t_s, d_arrived_inter, d_arrived_stop, v_max_stop, v_max_harddecel, a_max0, a_max1,
 *      a_des0, a_des1, vmax0, vmax1, l, delta_max, k_p_v, k_p_delta, g] */
static double dist_q(const twod_t* a, int64_t k, int idx) {
    const double* d = a->destq + (k * a->qcap + idx) * 3;
    return hypot(d[0] - a->s[k * 5], d[1] - a->s[k * 5 + 1]);
}
static void update_destination(twod_t* a, int64_t k, const double* p) {
    const double dnext = dist_q(a, k, a->ptr[k]);
    if (a->znav[k] & 6) return;
    if (dnext <= p[1]) a->ptr[k] = a->ptr[k] + 1 < a->qlen[k] - 1 ? a->ptr[k] + 1 : a->qlen[k] - 1;
    if (a->ptr[k] < a->qlen[k] - 1 && dist_q(a, k, a->ptr[k] + 1) < dnext) a->ptr[k] += 1;
}
static double update_nav_state(twod_t* a, int64_t k, const double* p, int stop, double* dd_out) {
    const double kk = 1.5, v = a->s[k * 5 + 3], vhd = p[4];
    const int z = a->znav[k], z0 = z & 1, z1 = (z >> 1) & 1, z2 = (z >> 2) & 1;
    double d0, d1;
    if (z0) { d0 = 0.5 * (vhd * vhd - v * v) / p[7]; d1 = 0.5 * -(vhd * vhd) / p[5]; }
    else { d0 = a->znavp[k * 3 + 1]; d1 = a->znavp[k * 3 + 2]; }
    const double dd = dist_q(a, k, a->ptr[k]);
    const int x0 = stop != 0, x1 = dd <= kk * (d0 + d1), x2 = dd <= p[2], x3 = v <= p[3];
    const int n0 = !x0 || (x0 && !x1 && ((z0 && !x2) || z1));
    const int n1 = x0 && ((z0 && ((!x2 && x1) || (x2 && !x3))) || (z1 && x1 && (!x2 || !x3)));
    const int n2 = x0 && (((z0 || z1) && x2 && x3) || z2);
    a->znav[k] = n0 | (n1 << 1) | (n2 << 2);
    if (z0 && n1) { a->znavp[k * 3] = v; a->znavp[k * 3 + 1] = d0; a->znavp[k * 3 + 2] = d1; }
    *dd_out = dd;
    const double* zp = a->znavp + k * 3;
    if (n0) return a->vd[k];
    if (n1) return dd < kk * zp[2] ? vhd / zp[2] * dd * 1 / kk : (zp[0] - vhd) / zp[1] * (dd - zp[2]) * 1 / kk + vhd;
    return 0.0;
}
static void dest_force_direct(twod_t* a, int64_t k, const double* p, double* f) {
    update_destination(a, k, p);
    const double* d = a->destq + (k * a->qcap + a->ptr[k]) * 3;
    double dd;
    const double vd = update_nav_state(a, k, p, d[2] != 0.0, &dd);
    if (dd > 0) { f[0] = -vd * (a->s[k * 5] - d[0]) / dd; f[1] = -vd * (a->s[k * 5 + 1] - d[1]) / dd; }
    else f[0] = f[1] = 0.0;
}
static void dest_force_twod(twod_t* a, int64_t k, const double* p, double* f) {
    update_destination(a, k, p);
    const double* d = a->destq + (k * a->qcap + a->ptr[k]) * 3;
    const int stop = d[2] != 0.0;
    double dd;
    const double vd = update_nav_state(a, k, p, stop, &dd);
    const double* s = a->s + k * 5;
    const int i = a->i[k];
    if (i == 0) { f[0] = vd * cos(s[2]); f[1] = vd * sin(s[2]); return; }
    if (a->znav[k] & 4) { f[0] = f[1] = 0.0; return; }
    const int last = a->ptr[k] + 1 >= a->qlen[k];
    double px[6], py[6];
    int m;
    if (!last) {
        px[0] = a->prev[k * 2]; py[0] = a->prev[k * 2 + 1];
        px[1] = s[0]; py[1] = s[1];
        int nd = a->qlen[k] - a->ptr[k];
        if (nd > 4) nd = 4;
        for (int j = 0; j < nd; ++j) {
            const double* q = a->destq + (k * a->qcap + a->ptr[k] + j) * 3;
            px[2 + j] = q[0]; py[2 + j] = q[1];
        }
        m = 2 + nd;
    } else {
        const int back = i < 100 ? i : 100;
        const int row = (a->hstep[k] - back) & 127;
        px[0] = a->hist[(k * 128 + row) * 2]; py[0] = a->hist[(k * 128 + row) * 2 + 1];
        px[1] = a->prev[k * 2]; py[1] = a->prev[k * 2 + 1];
        px[2] = s[0]; py[2] = s[1];
        px[3] = d[0]; py[3] = d[1];
        m = 4;
    }
    spline_t sp;
    if (!spline_fit(&sp, px, py, m)) { dest_force_direct(a, k, p, f); return; }
    int i_s = 1;
    double o[6];
    if (last) {
        double best = 0.0;
        for (int j = 0; j < 20; ++j) {
            spline_eval(&sp, j == 19 ? 1.0 : j * (1.0 / 19.0), o);
            const double d2 = (o[0] - s[0]) * (o[0] - s[0]) + (o[1] - s[1]) * (o[1] - s[1]);
            if (j == 0 || d2 < best) { best = d2; i_s = j; }
        }
    }
    const int i_p = i_s + (stop ? 5 : 3);
    if (i_p < 20) {
        double q[6];
        spline_eval(&sp, i_s == 19 ? 1.0 : i_s * (1.0 / 19.0), o);
        spline_eval(&sp, i_p == 19 ? 1.0 : i_p * (1.0 / 19.0), q);
        const double sp1 = sqrt(o[2] * o[2] + o[3] * o[3]);
        const double R = sp1 * sp1 * sp1 / fabs(o[2] * o[5] - o[3] * o[4]);
        double v = fmax(2.5, sqrt(10 * (TWO_PI / 360) * p[15] * R));
        v = fmin(v, vd);
        const double ex = q[0] - o[0], ey = q[1] - o[1], temp = v / sqrt(ex * ex + ey * ey);
        f[0] = temp * ex; f[1] = temp * ey;
    } else dest_force_direct(a, k, p, f);
}
static void control_move(twod_t* a, int64_t k, const double* p, double Fx, double Fy) {
    double* s = a->s + k * 5;
    const double th = atan2(Fy, Fx);
    double vF = sqrt(Fx * Fx + Fy * Fy);
    const double dd = dist_q(a, k, a->ptr[k]);
    if (dd < 3 && a->ptr[k] + 1 >= a->qlen[k]) vF = (vF / 3) * dd;
    const double target = angle_difference(s[2], th);
    const double ddelta = angle_difference(s[4], target);
    const double acc = clampd(p[13] * (vF - s[3]), p[5], p[6]);
    const double od = p[14] * ddelta, ts = p[0];
    double delta = limit_angle(s[4] + ts * od);
    double v = s[3] + ts * acc;
    delta = clampd(delta, -p[12], p[12]);
    v = clampd(v, p[9], p[10]);
    const double psi = limit_angle(s[2] + ts * v * tan(delta) / p[11]);
    s[1] += ts * v * sin(psi);
    s[0] += ts * v * cos(psi);
    s[2] = psi; s[3] = v; s[4] = delta;
}

/* One step of the sub-group: destination force, clip of frep (already summed over ALL sources),
 * total force, control + move, history.  n_total: road users in the whole crowd. */
void csf_c_twod_step(int64_t ns, int64_t qcap, double* s, int32_t* i, int32_t* ptr, const int32_t* qlen,
                     const double* destq, int32_t* znav, double* znavp, const double* vd, double* prev,
                     double* hist, int32_t* hstep, const double* p, int64_t n_total, const double* frep,
                     double* force) {
    twod_t a = {ns, qcap, s, i, ptr, znav, qlen, destq, znavp, vd, prev, hist, hstep};
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < ns; ++k) {
        double fd[2];
        dest_force_twod(&a, k, p, fd);
        double frx = 0.0, fry = 0.0;
        if (n_total > 1) {
            frx = frep[k * 2]; fry = frep[k * 2 + 1];
            const double rin = hypot(frx, fry), r = hypot(fd[0], fd[1]);
            if (rin > r) { frx = frx * r / rin; fry = fry * r / rin; }
        }
        const double Fx = frx + fd[0], Fy = fry + fd[1];
        force[k * 2] = Fx; force[k * 2 + 1] = Fy;
        const double ox = s[k * 5], oy = s[k * 5 + 1];
        if (znav[k] & 4) { s[k * 5 + 3] = 0.0; s[k * 5 + 4] = 0.0; }
        else control_move(&a, k, p, Fx, Fy);
        i[k] = (i[k] + 1) % 3000;
        prev[k * 2] = ox; prev[k * 2 + 1] = oy;
        hstep[k] += 1;
        const int row = hstep[k] & 127;
        hist[(k * 128 + row) * 2] = s[k * 5];
        hist[(k * 128 + row) * 2 + 1] = s[k * 5 + 1];
    }
}
