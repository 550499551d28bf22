/*
 * fgd_mirror.c -- CPU "mirror" oracle for the batched FGD hot path.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  Never by the product.
 *
 * What it is: a plain-C FP32 restatement of the reference algorithm
 *   trajectory.py:63-65,81-137,183-297   (evaluate, costs, gradients, checks)
 *   robot.py:29-36,75-113                (fk, jacobian, predicates)
 *   environment.py:32-58                 (obstacle cost / gradient)
 *   optimizer_BLS.py:126-213             (jit loop semantics)
 *   optimizer_GD.py:68-97,172-232        (single-level and dual loops)
 *   trajectory.py:73-78                  (initTrajectory, rank-2 form: mirror_init)
 * with every floating-point operation written out explicitly (fmaf where a
 * fused multiply-add is meant, separate * and + elsewhere; compiled with
 * -ffp-contract=off) and every reduction performed in one fixed, documented
 * order ("32-lane tree order", see tree_sum()).  The reference leaves its
 * summation order to XLA:CPU, so any fixed order is an equally valid reading;
 * fixing it lets the CUDA kernels be compared BIT-EXACTLY with this file in
 * their strict-math mode.  The looser link (this file <-> the line-by-line
 * NumPy restatement oracle/fgd_numpy.py <-> the reference's golden files) is
 * checked on the CPU in tests/test_oracle_*.py.
 *
 * Deliberate, documented re-associations w.r.t. the NumPy restatement (all
 * within 1-2 ulp per operation; tolerances in tests/test_oracle_mirror.py):
 *   - the obstacle term works with m = 2 den = 1 + |f-o|^2 (fma(dy,dy,fma(dx,dx,1))),
 *     accumulates sum_o 1/m and sum_o d/m^2 (reciprocal, not division) and applies
 *     the factors 0.8*2 / -0.8*4 once per time sample (the powers of two are exact)
 *                                                  (environment.py:43,57)
 *   - (q-mean)/std, /std^2, v/vmax, /T in the limit penalties multiply by
 *     host-rounded reciprocals                     (trajectory.py:217,232,247,260)
 *   - the gradient is normalised by multiplying with 1/||g|| instead of dividing
 *                                                  (optimizer_BLS.py:165)
 *   - alpha_norm = sum_t (sum_a g[t,a]) * (sum_b n[t,b]), algebraically equal
 *     to sum(g.T @ n)                              (optimizer_BLS.py:166)
 *   - sin/cos: Cody-Waite reduction + cephes minimax polynomials written out
 *     below (<= 2 ulp), instead of the platform libm; the quadrant comes from a
 *     magic-number rounding fma(x, 2/pi, 1.5 * 2^23) (round to nearest even).
 *   - trajectories of T <= 64 samples facing >= 64 obstacles sum the potential of a
 *     sample as TWO chains (share_split(): the CUDA side runs the second chain, whole
 *     blocks of four obstacles, on the lanes of the warp that own no sample); checked
 *     against the FP64 NumPy oracle in tests/test_oracle_mirror.py.
 *
 * Extension (cfg.whole_arm, not in the reference's code; its blog, DevBlog-Theme/
 * blog-post.html:505-513, names it): cost_v[t] = sum_j costmap(fk_j(q_t)) over the three
 * joint positions (robot.py:39-72), summed with explicit fmas; gradient through the
 * per-joint Jacobians.  Pinned by oracle/fgd_numpy.py (finite-difference checked in FP64).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAX_T 512
#define FS 8   /* floats per trajectory state  */
#define IS 8   /* ints per trajectory state    */

typedef struct {
    int T, n_obs, max_inner, max_outer, max_bls, cvdl, mode, strict;
    float lam_sg0, lam_jl0, lam_inc, lam_max, lam_reg;
    float eps_loop, eps_pos, eps_vel;
    float bls_lr0, bls_alpha, bls_bp, bls_bm;
    float safety, qmax, qmin, vmax;
    float link[3];
    float J[9];
    float gd_lr[16];
    int whole_arm;      /* 1: obstacle cost summed over all joint positions fk_1..fk_3 (blog-post.html:505-513) */
} MirrorCfg;

/* state layout shared with the CUDA side (include/fgd_b200.h) */
enum { F_LAM_SG = 0, F_LAM_JL, F_LR, F_LOSS, F_TOC, F_LAST_NEW_LOSS, F_R6, F_R7 };
enum { I_STATUS = 0, I_OUTER, I_INNER, I_INNER_TOTAL, I_CAND_EVALS, I_ACCEPTS, I_FULFILLED, I_HASH };
enum { ST_FRESH = 0, ST_ACTIVE = 1, ST_DONE = 2 };

typedef struct {
    /* derived constants, all rounded once to FP32 exactly like the device side */
    float oml, inv_T, w_avg, mean_q, inv_std, inv_std2, inv_vmax, inv_vmax2, q_hi, q_lo, v_hi, fT;
} Derived;

static void derive(const MirrorCfg *c, Derived *d)
{
    d->fT = (float)c->T;
    d->oml = 1.0f - c->lam_max;
    d->inv_T = 1.0f / d->fT;
    d->w_avg = d->oml * d->inv_T;
    d->mean_q = 0.5f * (c->qmax + c->qmin);                 /* trajectory.py:31 */
    float std_q = 0.5f * (c->qmax - d->mean_q);             /* trajectory.py:32 */
    d->inv_std = 1.0f / std_q;
    d->inv_std2 = 1.0f / (std_q * std_q);
    d->inv_vmax = 1.0f / c->vmax;
    d->inv_vmax2 = 1.0f / (c->vmax * c->vmax);
    d->q_hi = c->safety * c->qmax;                          /* trajectory.py:221 */
    d->q_lo = c->safety * c->qmin;                          /* trajectory.py:222 */
    d->v_hi = c->safety * c->vmax;                          /* trajectory.py:251 */
}

/* ---- sin/cos ------------------------------------------------------------ */
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static void mirror_sincos(float x, float *s_out, float *c_out)
{
    /* magic-number rounding: the low mantissa bits of t = fma(x, 2/pi, 1.5 * 2^23) are round-to-nearest-even(x * 2/pi)
     * in two's complement, j = t - 1.5 * 2^23 is that integer as a float (same sequence as the CUDA side) */
    const float t = fmaf(x, 6.366197467e-01f, 12582912.0f);
    const float j = t + -12582912.0f;
    float r = fmaf(j, -1.570796371e+00f, x);           /* Cody-Waite: pi/2 = C1+C2+C3 */
    r = fmaf(j, 4.371138829e-08f, r);
    r = fmaf(j, 1.715124510e-15f, r);
    float s = r * r;
    /* sin(r), |r| <= pi/4 */
    float ps = fmaf(s, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(ps, s, -1.6666654611e-1f);
    float sn = fmaf(ps * s, r, r);
    /* cos(r) */
    float pc = fmaf(s, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(pc, s, 4.166664568298827e-2f);
    float cs = fmaf(pc * s, s, fmaf(-0.5f, s, 1.0f));
    const uint32_t n = f2u(t);                          /* quadrant = n & 3 */
    const float s0 = (n & 1u) ? cs : sn, c0 = (n & 1u) ? sn : cs;
    *s_out = u2f(f2u(s0) ^ ((n << 30) & 0x80000000u));          /* -sin for n & 3 >= 2 */
    *c_out = u2f(f2u(c0) ^ (((n + 1u) << 30) & 0x80000000u));   /* -cos for n & 3 == 1, 2 */
}

/* ---- fixed-order reductions -------------------------------------------- */
/* A trajectory is owned by a team of WPT warps (1, 2 or 4 for T <= 64, 128, 256); team thread i
 * owns the 2 adjacent rows t = 2i, 2i+1.  Sum over t: thread partial = sequential sum over its
 * rows (from +0), then an xor butterfly 16, 8, 4, 2, 1 inside each warp, then the warp partials
 * are combined as p0 + p1 (WPT = 2) or (p0 + p1) + (p2 + p3) (WPT = 4).  Same mapping as the
 * CUDA side (csrc/fgd_device.cuh).                                                         */
static int warps_per_trajectory(int T) { return T <= 64 ? 1 : (T <= 128 ? 2 : (T <= 256 ? 4 : 8)); }

/* Obstacle sums of single-warp teams with many obstacles are TWO chains per sample, each from zero, added at the end
 * (csrc/fgd_device.cuh, share_split: chain B runs on the lanes of the warp that own no sample, in whole blocks of four):
 *     chain A = [0, S) ++ [n_obs - rem, n_obs),  chain B = [S, n_obs - rem),  rem = (n_obs - S) mod 4,  each ascending.
 * share_split() returns S; 0 = one chain. */
static int share_split(int T, int n_obs, int whole_arm)
{
    if (T > 64 || whole_arm || n_obs < 64) return 0;
    const int n_act = (T + 1) / 2, n_help = 32 - n_act;
    if (n_help <= 0) return 0;
    const int k = (n_act + n_help - 1) / n_help;
    const int lseg = 4 * ((n_obs + 4 * (k + 1) - 1) / (4 * (k + 1)));
    const int S = k * lseg;
    return S < n_obs ? S : 0;
}

static float tree_sum(const float *x, int T)
{
    const int WPT = warps_per_trajectory(T);
    float wp[8];
    for (int w = 0; w < WPT; ++w) {
        float p[32];
        for (int l = 0; l < 32; ++l) {
            float a = 0.0f;
            for (int r = 0; r < 2; ++r) { const int t = 2 * (32 * w + l) + r; if (t < T) a = a + x[t]; }
            p[l] = a;
        }
        for (int off = 16; off >= 1; off >>= 1) {
            float q[32];
            for (int l = 0; l < 32; ++l) q[l] = p[l] + p[l ^ off];
            memcpy(p, q, sizeof(p));
        }
        wp[w] = p[0];
    }
    for (int n = WPT; n > 1; n >>= 1)
        for (int w = 0; w < n / 2; ++w) wp[w] = wp[2 * w] + wp[2 * w + 1];
    return wp[0];
}

static inline float ss3(float a, float b, float c) { return fmaf(c, c, fmaf(b, b, a * a)); }

/* ---- one evaluation: q,v -> loss, G_q, G_v, flags ---------------------- */
typedef struct {
    float loss, toc;
    int fulfilled;
} EvalOut;

static void contract(const float *Mt, const float *X, float sign, int T, float *Y)
{
    /* Y[i][a] = sum_k fma(M[i][k], sign*X[k][a], acc), k ascending for every (i,a).
     * Mt is M transposed (Mt[k][i] = M[i][k]) so the i loop is unit-stride and
     * vectorises; each accumulator still sees its k terms strictly in order. */
    float a0[MAX_T], a1[MAX_T], a2[MAX_T];
    for (int i = 0; i < T; ++i) a0[i] = a1[i] = a2[i] = 0.0f;
    for (int k = 0; k < T; ++k) {
        const float x0 = sign * X[k * 3], x1 = sign * X[k * 3 + 1], x2 = sign * X[k * 3 + 2];
        const float *m = Mt + (size_t)k * T;
        for (int i = 0; i < T; ++i) {
            a0[i] = fmaf(m[i], x0, a0[i]);
            a1[i] = fmaf(m[i], x1, a1[i]);
            a2[i] = fmaf(m[i], x2, a2[i]);
        }
    }
    for (int i = 0; i < T; ++i) { Y[i * 3] = a0[i]; Y[i * 3 + 1] = a1[i]; Y[i * 3 + 2] = a2[i]; }
}

static void evaluate_point(const MirrorCfg *c, const Derived *d, const float *K, const float *dK,
                           const float *obs, const float *alpha, const float *start, const float *goal,
                           float lam_sg, float lam_jl, float *q, float *v, float *Gq, float *Gv, EvalOut *out)
{
    const int T = c->T;
    float qr[MAX_T * 3], vr[MAX_T * 3];
    float costv[MAX_T], gx[MAX_T][3], gy[MAX_T][3], jp_row[MAX_T], jv_row[MAX_T];
    const int j_lo = c->whole_arm ? 0 : 2;                   /* joints whose position is charged */
    float sn[MAX_T][3], cs[MAX_T][3];
    contract(K, alpha, 1.0f, T, qr);                         /* trajectory.py:273 */
    contract(dK, alpha, 1.0f, T, vr);                        /* trajectory.py:274 */
    const float *J = c->J;
    int q_ok = 1, v_ok = 1;
    for (int t = 0; t < T; ++t) {
        for (int b = 0; b < 3; ++b) {                        /* (M@alpha)@J  trajectory.py:65 */
            q[t * 3 + b] = fmaf(qr[t * 3 + 2], J[6 + b], fmaf(qr[t * 3 + 1], J[3 + b], qr[t * 3] * J[b]));
            v[t * 3 + b] = fmaf(vr[t * 3 + 2], J[6 + b], fmaf(vr[t * 3 + 1], J[3 + b], vr[t * 3] * J[b]));
        }
        const float *qt = q + t * 3, *vt = v + t * 3;
        float c1 = qt[0], c2 = c1 + qt[1], c3 = c2 + qt[2]; /* robot.py:32 */
        mirror_sincos(c1, &sn[t][0], &cs[t][0]);
        mirror_sincos(c2, &sn[t][1], &cs[t][1]);
        mirror_sincos(c3, &sn[t][2], &cs[t][2]);
        float px[3], py[3], cj[3], srj[3];                           /* joint positions fk_1, fk_2, fk_3 = fk   robot.py:33-34, 39-72 */
        px[0] = c->link[0] * cs[t][0]; px[1] = fmaf(c->link[1], cs[t][1], px[0]); px[2] = fmaf(c->link[2], cs[t][2], px[1]);
        py[0] = c->link[0] * sn[t][0]; py[1] = fmaf(c->link[1], sn[t][1], py[0]); py[2] = fmaf(c->link[2], sn[t][2], py[1]);
        const int split = share_split(T, c->n_obs, c->whole_arm);
        const int tail_end = split > 0 ? c->n_obs - ((c->n_obs - split) & 3) : c->n_obs;
        for (int j = j_lo; j < 3; ++j) {
            float acc[2][3] = { { 0.0f, 0.0f, 0.0f }, { 0.0f, 0.0f, 0.0f } };   /* chain A, chain B: sr, sx, sy */
            for (int pass = 0; pass < (split > 0 ? 2 : 1); ++pass) {
                for (int o = 0; o < c->n_obs; ++o) {         /* environment.py:46-58 */
                    const int in_b = split > 0 && o >= split && o < tail_end;
                    if (in_b != pass) continue;              /* pass 0: chain A (ascending, the remainder last), pass 1: chain B */
                    float *a = acc[pass];
                    float dx = px[j] - obs[2 * o], dy = py[j] - obs[2 * o + 1];
                    float m = fmaf(dy, dy, fmaf(dx, dx, 1.0f));  /* 2 den = 1 + |f-o|^2  */
                    float r = 1.0f / m;
                    a[0] = a[0] + r;                         /* sum 1/(2 den)        */
                    float r2 = r * r;
                    a[1] = fmaf(r2, dx, a[1]);               /* sum d/(2 den)^2      */
                    a[2] = fmaf(r2, dy, a[2]);
                }
            }
            float sr = acc[0][0], sx = acc[0][1], sy = acc[0][2];
            if (split > 0) { sr = sr + acc[1][0]; sx = sx + acc[1][1]; sy = sy + acc[1][2]; }
            srj[j] = sr; cj[j] = 1.6f * sr; gx[t][j] = -3.2f * sx; gy[t][j] = -3.2f * sy;
        }
        /* whole arm: c_1 + c_2 + c_3 as fma(1.6, s_3, fma(1.6, s_2, 1.6 * s_1)) */
        costv[t] = c->whole_arm ? fmaf(1.6f, srj[2], fmaf(1.6f, srj[1], cj[0])) : cj[2];
        /* joint-limit penalties  trajectory.py:215-268 */
        float ep = 0.0f, ev = 0.0f;
        float e3[3], f3[3];
        for (int b = 0; b < 3; ++b) {
            float u = (qt[b] - d->mean_q) * d->inv_std;
            int m = c->cvdl ? (qt[b] > d->q_hi || qt[b] < d->q_lo) : 1;
            e3[b] = m ? 0.5f * (u * u) : 0.0f;
            float w = vt[b] * d->inv_vmax;
            int mv = c->cvdl ? (fabsf(vt[b]) > d->v_hi) : 1;
            f3[b] = mv ? 0.5f * (w * w) : 0.0f;
            if (!(qt[b] <= c->qmax) || !(qt[b] >= c->qmin)) q_ok = 0;     /* robot.py:104-108 */
            if (!(fabsf(vt[b]) <= c->vmax)) v_ok = 0;                      /* robot.py:111-113 */
        }
        ep = (e3[0] + e3[1]) + e3[2];
        ev = (f3[0] + f3[1]) + f3[2];
        jp_row[t] = ep; jv_row[t] = ev;
    }
    /* max / argmax(first) / mean   trajectory.py:81-110 */
    float maxc = costv[0]; int amax = 0;
    for (int t = 1; t < T; ++t) if (costv[t] > maxc) { maxc = costv[t]; amax = t; }
    float sumc = tree_sum(costv, T);
    float avg = sumc / d->fT;
    float toc = fmaf(c->lam_max, maxc, d->oml * avg);
    /* start/goal terms  trajectory.py:183-212 */
    const float *q0 = q, *qT = q + (T - 1) * 3, *v0 = v, *vT = v + (T - 1) * 3;
    float d0[3] = { q0[0] - start[0], q0[1] - start[1], q0[2] - start[2] };
    float dT[3] = { qT[0] - goal[0], qT[1] - goal[1], qT[2] - goal[2] };
    float ssp0 = ss3(d0[0], d0[1], d0[2]), sspT = ss3(dT[0], dT[1], dT[2]);
    float ssv0 = ss3(v0[0], v0[1], v0[2]), ssvT = ss3(vT[0], vT[1], vT[2]);
    float sg = (0.5f * ssp0 + 0.5f * sspT) + (0.5f * ssv0 + 0.5f * ssvT);
    float jl = tree_sum(jp_row, T) / d->fT + tree_sum(jv_row, T) / d->fT;
    out->toc = toc;
    out->loss = fmaf(lam_jl, jl, fmaf(lam_sg, sg, toc));      /* trajectory.py:281 */
    out->fulfilled = (sqrtf(ssp0) < c->eps_pos) && (sqrtf(sspT) < c->eps_pos) &&
                     (sqrtf(ssv0) < c->eps_vel) && (sqrtf(ssvT) < c->eps_vel) && q_ok && v_ok;   /* trajectory.py:129-137 */
    /* gradients w.r.t. q and v rows */
    for (int t = 0; t < T; ++t) {
        const float *qt = q + t * 3, *vt = v + t * 3;
        float wt = (t == amax) ? (c->lam_max + d->w_avg) : d->w_avg;      /* trajectory.py:100-105 */
        float cgx = wt * gx[t][2], cgy = wt * gy[t][2];
        float xs[3], ys[3];
        for (int k = 0; k < 3; ++k) { xs[k] = -(c->link[k] * sn[t][k]); ys[k] = c->link[k] * cs[t][k]; }   /* robot.py:80,83 */
        float Sx = (xs[0] + xs[1]) + xs[2], Sy = (ys[0] + ys[1]) + ys[2];
        float Cx[3] = { xs[0], xs[0] + xs[1], (xs[0] + xs[1]) + xs[2] };
        float Cy[3] = { ys[0], ys[0] + ys[1], (ys[0] + ys[1]) + ys[2] };
        for (int k = 0; k < 3; ++k) {
            float Jx = (xs[k] + Sx) - Cx[k];                               /* robot.py:81 */
            float Jy = (ys[k] + Sy) - Cy[k];                               /* robot.py:84 */
            float tg = fmaf(cgy, Jy, cgx * Jx);                            /* trajectory.py:125 */
            if (c->whole_arm) {                                            /* + joints 2 and 1: J_j[k] = sum_{m=k..j} */
                if (k <= 1) {
                    float J2x = (k == 0) ? xs[0] + xs[1] : xs[1], J2y = (k == 0) ? ys[0] + ys[1] : ys[1];
                    tg = fmaf(wt * gy[t][1], J2y, fmaf(wt * gx[t][1], J2x, tg));
                }
                if (k == 0) tg = fmaf(wt * gy[t][0], ys[0], fmaf(wt * gx[t][0], xs[0], tg));
            }
            float sgp = (t == 0) ? d0[k] : ((t == T - 1) ? dT[k] : 0.0f);
            float sgv = (t == 0 || t == T - 1) ? vt[k] : 0.0f;
            int m = c->cvdl ? (qt[k] > d->q_hi || qt[k] < d->q_lo) : 1;
            float jpg = m ? ((qt[k] - d->mean_q) * d->inv_std2) * d->inv_T : 0.0f;
            int mv = c->cvdl ? (fabsf(vt[k]) > d->v_hi) : 1;
            float jvg = mv ? (vt[k] * d->inv_vmax2) * d->inv_T : 0.0f;
            Gq[t * 3 + k] = fmaf(lam_jl, jpg, fmaf(lam_sg, sgp, tg));     /* trajectory.py:295 */
            Gv[t * 3 + k] = fmaf(lam_jl, jvg, lam_sg * sgv);
        }
    }
}

/* alpha-gradient  (K^T Gq + dK^T Gv) J^T ; K symmetric, dK antisymmetric      trajectory.py:295 */
static void backward(const MirrorCfg *c, const float *K, const float *dK, const float *Gq, const float *Gv, float *g)
{
    const int T = c->T;
    float y1[MAX_T * 3], y2[MAX_T * 3];
    contract(K, Gq, 1.0f, T, y1);
    contract(dK, Gv, -1.0f, T, y2);      /* dK^T = -dK exactly in FP32 */
    const float *J = c->J;
    for (int t = 0; t < T; ++t) {
        float r0 = y1[t * 3] + y2[t * 3], r1 = y1[t * 3 + 1] + y2[t * 3 + 1], r2 = y1[t * 3 + 2] + y2[t * 3 + 2];
        for (int b = 0; b < 3; ++b) g[t * 3 + b] = fmaf(r2, J[b * 3 + 2], fmaf(r1, J[b * 3 + 1], r0 * J[b * 3]));
    }
}

static float *transpose(const float *M, int T)
{
    float *t = (float *)malloc(sizeof(float) * T * T);
    for (int i = 0; i < T; ++i) for (int k = 0; k < T; ++k) t[k * T + i] = M[i * T + k];
    return t;
}

int mirror_eval(const MirrorCfg *c, const float *K_in, const float *dK_in, const float *obs, int B,
                const float *alpha, const float *start, const float *goal, float lam_sg, float lam_jl,
                float *loss, float *toc, float *grad, float *q_out, float *v_out, int *fulfilled)
{
    if (c->T > MAX_T || c->T < 2) return 1;
    Derived d; derive(c, &d);
    const int T = c->T;
    float *Kt = transpose(K_in, T), *dKt = transpose(dK_in, T);
    const float *K = Kt, *dK = dKt;
#pragma omp parallel for schedule(dynamic, 4)
    for (int b = 0; b < B; ++b) {
        float q[MAX_T * 3], v[MAX_T * 3], Gq[MAX_T * 3], Gv[MAX_T * 3], g[MAX_T * 3];
        EvalOut eo;
        evaluate_point(c, &d, K, dK, obs, alpha + (size_t)b * T * 3, start + b * 3, goal + b * 3, lam_sg, lam_jl, q, v, Gq, Gv, &eo);
        backward(c, K, dK, Gq, Gv, g);
        if (loss) loss[b] = eo.loss;
        if (toc) toc[b] = eo.toc;
        if (fulfilled) fulfilled[b] = eo.fulfilled;
        if (grad) memcpy(grad + (size_t)b * T * 3, g, sizeof(float) * T * 3);
        if (q_out) memcpy(q_out + (size_t)b * T * 3, q, sizeof(float) * T * 3);
        if (v_out) memcpy(v_out + (size_t)b * T * 3, v, sizeof(float) * T * 3);
    }
    free(Kt); free(dKt);
    return 0;
}

static inline void hash_step(int *h, int code) { *h = (int)((uint32_t)(*h) * 1000003u + (uint32_t)code); }

/* One trajectory, up to `budget` inner iterations (budget < 0: unlimited). */
static void optimize_one(const MirrorCfg *c, const Derived *d, const float *K, const float *dK, const float *obs,
                         float *alpha, const float *start, const float *goal, float *fs, int *is, int budget)
{
    const int T = c->T, n = T * 3;
    if (is[I_STATUS] == ST_DONE) return;
    if (is[I_STATUS] == ST_FRESH) {
        fs[F_LAM_SG] = c->lam_sg0; fs[F_LAM_JL] = c->lam_jl0;
        fs[F_LR] = (c->mode == 0) ? c->bls_lr0 : c->gd_lr[0];
        is[I_OUTER] = 0; is[I_INNER] = 0; is[I_INNER_TOTAL] = 0; is[I_CAND_EVALS] = 0; is[I_ACCEPTS] = 0;
        is[I_FULFILLED] = 0; is[I_HASH] = 0;
        is[I_STATUS] = ST_ACTIVE;
    }
    float q[MAX_T * 3], v[MAX_T * 3], Gq[MAX_T * 3], Gv[MAX_T * 3], g[MAX_T * 3], nh[MAX_T * 3];
    float cand[MAX_T * 3], Gqc[MAX_T * 3], Gvc[MAX_T * 3], rowa[MAX_T], rowb[MAX_T];
    int done_iters = 0;
    for (;;) {
        float lam_sg = fs[F_LAM_SG], lam_jl = fs[F_LAM_JL];
        if (c->mode == 1) fs[F_LR] = c->gd_lr[is[I_OUTER]];             /* optimizer_GD.py:209 */
        EvalOut e0;
        evaluate_point(c, d, K, dK, obs, alpha, start, goal, lam_sg, lam_jl, q, v, Gq, Gv, &e0);   /* :163 / GD :210 */
        float loss = e0.loss; int ful = e0.fulfilled; float toc = e0.toc;
        int minimized = 0;
        while (is[I_INNER] < c->max_inner && !minimized) {
            if (budget >= 0 && done_iters == budget) { fs[F_LOSS] = loss; fs[F_TOC] = toc; is[I_FULFILLED] = ful; return; }
            ++done_iters; ++is[I_INNER_TOTAL];
            backward(c, K, dK, Gq, Gv, g);                               /* :164 */
            float lr = fs[F_LR];
            if (c->mode == 0) {
                for (int t = 0; t < T; ++t) rowa[t] = ss3(g[t * 3], g[t * 3 + 1], g[t * 3 + 2]);
                float scale = 1.0f / sqrtf(tree_sum(rowa, T));
                for (int i = 0; i < n; ++i) nh[i] = g[i] * scale;        /* :165 (reciprocal, then multiply) */
                for (int t = 0; t < T; ++t)
                    rowb[t] = ((g[t * 3] + g[t * 3 + 1]) + g[t * 3 + 2]) * ((nh[t * 3] + nh[t * 3 + 1]) + nh[t * 3 + 2]);
                float alpha_norm = tree_sum(rowb, T);                    /* :166 */
                float new_loss = loss; int accepted = 0, j = 0;
                for (j = 0; j < c->max_bls; ++j) {                       /* :131-150 */
                    float c1 = 1.0f - c->lam_reg * lr;
                    for (int i = 0; i < n; ++i) cand[i] = fmaf(c1, alpha[i], -(lr * nh[i]));
                    EvalOut ec;
                    evaluate_point(c, d, K, dK, obs, cand, start, goal, lam_sg, lam_jl, q, v, Gqc, Gvc, &ec);
                    ++is[I_CAND_EVALS];
                    float req = loss - (c->bls_alpha * lr) * alpha_norm;
                    if (ec.loss > req) { lr = lr * c->bls_bm; hash_step(&is[I_HASH], 1); }
                    else {
                        memcpy(alpha, cand, sizeof(float) * n); memcpy(Gq, Gqc, sizeof(float) * n); memcpy(Gv, Gvc, sizeof(float) * n);
                        new_loss = ec.loss; ful = ec.fulfilled; toc = ec.toc;
                        lr = lr * c->bls_bp; accepted = 1; ++is[I_ACCEPTS]; hash_step(&is[I_HASH], 2);
                        break;
                    }
                }
                fs[F_LR] = lr;
                fs[F_LAST_NEW_LOSS] = new_loss;
                minimized = (loss - new_loss < c->eps_loop);             /* :178 */
                if (accepted) loss = new_loss;
                if (!minimized) ++is[I_INNER];
                else hash_step(&is[I_HASH], 3);
            } else {
                float c1 = 1.0f - c->lam_reg * lr;
                for (int i = 0; i < n; ++i) cand[i] = fmaf(c1, alpha[i], -(lr * g[i]));    /* GD :185 */
                EvalOut ec;
                evaluate_point(c, d, K, dK, obs, cand, start, goal, lam_sg, lam_jl, q, v, Gqc, Gvc, &ec);
                ++is[I_CAND_EVALS];
                fs[F_LAST_NEW_LOSS] = ec.loss;
                if (loss - ec.loss < c->eps_loop) { minimized = 1; hash_step(&is[I_HASH], 3); }   /* GD :194 */
                else {
                    memcpy(alpha, cand, sizeof(float) * n); memcpy(Gq, Gqc, sizeof(float) * n); memcpy(Gv, Gvc, sizeof(float) * n);
                    loss = ec.loss; ful = ec.fulfilled; toc = ec.toc;
                    ++is[I_ACCEPTS]; ++is[I_INNER]; hash_step(&is[I_HASH], 2);
                }
            }
        }
        fs[F_LOSS] = loss; fs[F_TOC] = toc; is[I_FULFILLED] = ful;
        int dual = (c->mode == 0) || (c->max_outer > 1);
        if (!dual || ful) { is[I_STATUS] = ST_DONE; return; }            /* :196-205 */
        fs[F_LAM_SG] = fs[F_LAM_SG] * c->lam_inc; fs[F_LAM_JL] = fs[F_LAM_JL] * c->lam_inc;
        ++is[I_OUTER]; hash_step(&is[I_HASH], 4);
        if (is[I_OUTER] >= c->max_outer) { is[I_STATUS] = ST_DONE; return; }
        is[I_INNER] = 0;
        if (c->mode == 0) fs[F_LR] = c->bls_lr0;                         /* :193 */
    }
}

int mirror_optimize(const MirrorCfg *c, const float *K_in, const float *dK_in, const float *obs, int B,
                    float *alpha, const float *start, const float *goal, float *fstate, int *istate,
                    int budget, int nthreads)
{
    if (c->T > MAX_T || c->T < 2 || c->max_outer > 16) return 1;
    Derived d; derive(c, &d);
    const int T = c->T;
    float *Kt = transpose(K_in, T), *dKt = transpose(dK_in, T);
    const float *K = Kt, *dK = dKt;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b)
        optimize_one(c, &d, K, dK, obs, alpha + (size_t)b * T * 3, start + b * 3, goal + b * 3,
                     fstate + (size_t)b * FS, istate + (size_t)b * IS, budget);
    free(Kt); free(dKt);
    return 0;
}

/* Rank-2 restatement of Trajectory.initTrajectory (trajectory.py:73-78), the op order of
 * fgd_init_kernel:  s' = start J^-1, d' = (goal - start) J^-1,
 *   alpha[b][t][j] = fmaf(w[t], d'[j], u[t] * s'[j])   with u = K^-1 1, w = K^-1 c (inputs). */
int mirror_init(int T, int B, const float *u, const float *w, const float *jinv, const float *start, const float *goal,
                float *alpha)
{
    for (int b = 0; b < B; ++b) {
        const float s0 = start[b * 3], s1 = start[b * 3 + 1], s2 = start[b * 3 + 2];
        const float e0 = goal[b * 3] - s0, e1 = goal[b * 3 + 1] - s1, e2 = goal[b * 3 + 2] - s2;
        for (int j = 0; j < 3; ++j) {
            const float sj = fmaf(s2, jinv[6 + j], fmaf(s1, jinv[3 + j], s0 * jinv[j]));
            const float dj = fmaf(e2, jinv[6 + j], fmaf(e1, jinv[3 + j], e0 * jinv[j]));
            for (int t = 0; t < T; ++t) alpha[((size_t)b * T + t) * 3 + j] = fmaf(w[t], dj, u[t] * sj);
        }
    }
    return 0;
}

int mirror_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
