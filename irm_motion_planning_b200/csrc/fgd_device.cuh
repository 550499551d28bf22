// fgd_device.cuh -- device-side building blocks of the batched FGD iteration (sm_100a).
//
// Mapping: one warp runs S independent trajectories as
// autonomous state machines.  Lane L owns time samples t = L + 32 r (r < RPL):
// the RKHS contraction produces row t of q = K alpha J and v = dK alpha J in the
// lane that then does forward kinematics, the obstacle potential and the penalty
// terms for that sample -- no shared-memory round trip between the two.  K and dK
// are staged once per CTA in shared memory (transposed, so a warp reads 32
// consecutive floats), the operand rows (alpha' or the q/v-gradients) sit in
// per-warp shared buffers and are broadcast.  Reductions over t are lane-serial
// followed by a 5-level xor butterfly, the order the mirror oracle reproduces.
//
// Compiled with -fmad=false: every fused multiply-add is an explicit fmaf(), so
// the operation sequence is the documented one (bit-exact against the oracle in
// strict-math mode).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fgd {

constexpr unsigned FULL = 0xffffffffu;

enum Kind : int { K_IDLE = 0, K_EVAL0 = 1, K_CAND = 2, K_BACK = 3 };

struct DevParams {
    int T, TP, n_obs, max_inner, max_outer, max_bls, cvdl, mode, budget, B;
    float lam_sg0, lam_jl0, lam_inc, lam_max, lam_reg, eps_loop, eps_pos, eps_vel;
    float bls_lr0, bls_alpha, bls_bp, bls_bm;
    float qmax, qmin, vmax;
    float link[3];
    float J[9];
    float gd_lr[16];
    // derived on the host, rounded once to FP32 (same expressions as the oracle)
    float oml, inv_T, w_avg, mean_q, inv_std, inv_std2, inv_vmax, inv_vmax2, q_hi, q_lo, v_hi, fT;
    const float *Kt, *dKt;   // [T][TP]: Kt[k][i] = K[i][k], zero padded for i >= T
    const float *obs;        // [n_obs][2]
    float *alpha;            // [B][T][3]
    const float *start, *goal;
    float *fstate;
    int *istate;
    unsigned *queue;
};

struct EvalPtrs {
    float lam_sg, lam_jl;
    float *loss, *toc, *grad, *q, *v;
    int *fulfilled;
};

// warp-uniform per-trajectory scalars (shared memory, one per slot)
struct Slot {
    int traj, status, outer, inner, inner_total, cand_evals, accepts, ful, j, done_iters;
    unsigned hash;
    float lam_sg, lam_jl, lr, loss, toc, alpha_norm, last_new;
    float start[3], goal[3];
};

__device__ __forceinline__ float wsum(float v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = v + __shfl_xor_sync(FULL, v, o);
    return v;
}

__device__ __forceinline__ float ss3(float a, float b, float c) { return fmaf(c, c, fmaf(b, b, a * a)); }

template <bool STRICT>
__device__ __forceinline__ float rcp(float x)
{
    if constexpr (STRICT) {
        return __frcp_rn(x);
    } else {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
}

// sin and cos of one angle: Cody-Waite reduction by pi/2 (3 constants) and the
// cephes minimax polynomials on [-pi/4, pi/4]; <= 2 ulp for the angles a 3-joint
// arm with limits [-1, 2] rad produces.  Same operation sequence as the oracle.
__device__ __forceinline__ void sincos_cw(float x, float &S, float &C)
{
    const float j = rintf(x * 6.366197467e-01f);
    float r = fmaf(j, -1.570796371e+00f, x);
    r = fmaf(j, 4.371138829e-08f, r);
    r = fmaf(j, 1.715124510e-15f, r);
    const float s = r * r;
    float ps = fmaf(s, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(ps, s, -1.6666654611e-1f);
    const float sn = fmaf(ps * s, r, r);
    float pc = fmaf(s, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(pc, s, 4.166664568298827e-2f);
    const float cs = fmaf(pc * s, s, fmaf(-0.5f, s, 1.0f));
    const int n = ((int)j) & 3;
    S = (n & 1) ? cs : sn;
    C = (n & 1) ? sn : cs;
    if (n == 1 || n == 2) C = -C;
    if (n >= 2) S = -S;
}

// ---------------------------------------------------------------------------
// RKHS contraction for S trajectories of one warp:
//   y1[s][r][a] = sum_k K [t_r][k] * x1[s][k][a]      (trajectory.py:65 / :295)
//   y2[s][r][a] = sum_k dK[t_r][k] * x2[s][k][a]
// k ascending, one fmaf per term.  K rows come from shared memory (KS) or L2.
// ---------------------------------------------------------------------------
template <int RPL, int S, bool KS>
__device__ __forceinline__ void contract(const float *__restrict__ Kt, const float *__restrict__ dKt, int T, int TP, int lane,
                                         const float4 *const (&x1)[S], const float4 *const (&x2)[S],
                                         float (&y1)[S][RPL][3], float (&y2)[S][RPL][3])
{
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int r = 0; r < RPL; ++r)
#pragma unroll
            for (int a = 0; a < 3; ++a) { y1[s][r][a] = 0.0f; y2[s][r][a] = 0.0f; }

    const float *kp = Kt + lane, *dp = dKt + lane;
#pragma unroll 2
    for (int k = 0; k < T; ++k) {
        float kv[RPL], dv[RPL];
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            if constexpr (KS) { kv[r] = kp[32 * r]; dv[r] = dp[32 * r]; }
            else { kv[r] = __ldg(kp + 32 * r); dv[r] = __ldg(dp + 32 * r); }
        }
        kp += TP; dp += TP;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const float4 xa = x1[s][k];
            const float4 xb = x2[s][k];
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                y1[s][r][0] = fmaf(kv[r], xa.x, y1[s][r][0]);
                y1[s][r][1] = fmaf(kv[r], xa.y, y1[s][r][1]);
                y1[s][r][2] = fmaf(kv[r], xa.z, y1[s][r][2]);
                y2[s][r][0] = fmaf(dv[r], xb.x, y2[s][r][0]);
                y2[s][r][1] = fmaf(dv[r], xb.y, y2[s][r][1]);
                y2[s][r][2] = fmaf(dv[r], xb.z, y2[s][r][2]);
            }
        }
    }
}

// Per-lane rows kept between the cost phase and the gradient phase.
template <int RPL>
struct Rows {
    float q[RPL][3], v[RPL][3], sn[RPL][3], cs[RPL][3], gx[RPL], gy[RPL];
    float d0[3], dT[3];   // q[0]-start, q[T-1]-goal (meaningful in the owning lanes)
    int amax;
};

// ---------------------------------------------------------------------------
// Cost phase: compute_trajectory_cost + constraintsFulfilled for one trajectory
// whose raw contraction rows are yq (K alpha) and yv (dK alpha).
//   trajectory.py:271-281 (total), :81-88 (max/mean), :183-255 (penalties),
//   :129-137 + robot.py:90-113 (constraint predicates), robot.py:29-36 (fk),
//   environment.py:32-58 (obstacle potential and its (x,y)-gradient).
// ---------------------------------------------------------------------------
template <int RPL, bool STRICT>
__device__ __forceinline__ void cost_phase(const DevParams &p, const float2 *__restrict__ sObs, int lane,
                                           const float (&yq)[RPL][3], const float (&yv)[RPL][3],
                                           const float *start, const float *goal, float lam_sg, float lam_jl,
                                           Rows<RPL> &R, float &loss, float &toc, int &ful)
{
    const int T = p.T;
    float cost[RPL];
    float part_c = 0.0f, part_p = 0.0f, part_v = 0.0f;
    unsigned maxbits = 0u;
    bool lim_ok = true;
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int t = lane + 32 * r;
        const bool valid = t < T;
#pragma unroll
        for (int b = 0; b < 3; ++b) {   // (M @ alpha) @ J
            R.q[r][b] = fmaf(yq[r][2], p.J[6 + b], fmaf(yq[r][1], p.J[3 + b], yq[r][0] * p.J[b]));
            R.v[r][b] = fmaf(yv[r][2], p.J[6 + b], fmaf(yv[r][1], p.J[3 + b], yv[r][0] * p.J[b]));
        }
        const float c1 = R.q[r][0], c2 = c1 + R.q[r][1], c3 = c2 + R.q[r][2];
        sincos_cw(c1, R.sn[r][0], R.cs[r][0]);
        sincos_cw(c2, R.sn[r][1], R.cs[r][1]);
        sincos_cw(c3, R.sn[r][2], R.cs[r][2]);
        const float x = fmaf(p.link[2], R.cs[r][2], fmaf(p.link[1], R.cs[r][1], p.link[0] * R.cs[r][0]));
        const float y = fmaf(p.link[2], R.sn[r][2], fmaf(p.link[1], R.sn[r][1], p.link[0] * R.sn[r][0]));
        float c = 0.0f, ax = 0.0f, ay = 0.0f;
        const int n_obs = p.n_obs;
#pragma unroll 4
        for (int o = 0; o < n_obs; ++o) {
            const float2 ob = sObs[o];
            const float dx = x - ob.x, dy = y - ob.y;
            const float n = fmaf(dy, dy, dx * dx);
            const float den = fmaf(0.5f, n, 0.5f);
            const float rr = rcp<STRICT>(den);
            const float cc = 0.8f * rr;
            c = c + cc;
            const float w = -(cc * rr);
            ax = fmaf(w, dx, ax);
            ay = fmaf(w, dy, ay);
        }
        cost[r] = c; R.gx[r] = ax; R.gy[r] = ay;
        float e3[3], f3[3];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const float qb = R.q[r][b], vb = R.v[r][b];
            const float u = (qb - p.mean_q) * p.inv_std;
            const bool m = p.cvdl ? (qb > p.q_hi || qb < p.q_lo) : true;
            e3[b] = m ? 0.5f * (u * u) : 0.0f;
            const float w = vb * p.inv_vmax;
            const bool mv = p.cvdl ? (fabsf(vb) > p.v_hi) : true;
            f3[b] = mv ? 0.5f * (w * w) : 0.0f;
            if (valid && (!(qb <= p.qmax) || !(qb >= p.qmin) || !(fabsf(vb) <= p.vmax))) lim_ok = false;
        }
        if (valid) {
            part_c = part_c + c;
            part_p = part_p + ((e3[0] + e3[1]) + e3[2]);
            part_v = part_v + ((f3[0] + f3[1]) + f3[2]);
            maxbits = max(maxbits, __float_as_uint(c));   // c >= 0: uint order == float order
        }
    }
    // max / first argmax / mean over t
    const float maxc = __uint_as_float(__reduce_max_sync(FULL, maxbits));
    int cand = 0x7fffffff;
#pragma unroll
    for (int r = RPL - 1; r >= 0; --r) {
        const int t = lane + 32 * r;
        if (t < T && cost[r] == maxc) cand = t;
    }
    R.amax = __reduce_min_sync(FULL, cand);
    const float avg = wsum(part_c) / p.fT;
    toc = fmaf(p.lam_max, maxc, p.oml * avg);

    // start / goal rows
    const int lT = (T - 1) & 31, rT = (T - 1) >> 5;
    float qT[3] = {0.f, 0.f, 0.f}, vT[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < RPL; ++r)
        if (r == rT) {
#pragma unroll
            for (int b = 0; b < 3; ++b) { qT[b] = R.q[r][b]; vT[b] = R.v[r][b]; }
        }
#pragma unroll
    for (int b = 0; b < 3; ++b) { R.d0[b] = R.q[0][b] - start[b]; R.dT[b] = qT[b] - goal[b]; }
    const float ssp0 = __shfl_sync(FULL, ss3(R.d0[0], R.d0[1], R.d0[2]), 0);
    const float ssv0 = __shfl_sync(FULL, ss3(R.v[0][0], R.v[0][1], R.v[0][2]), 0);
    const float sspT = __shfl_sync(FULL, ss3(R.dT[0], R.dT[1], R.dT[2]), lT);
    const float ssvT = __shfl_sync(FULL, ss3(vT[0], vT[1], vT[2]), lT);
    const float sg = (0.5f * ssp0 + 0.5f * sspT) + (0.5f * ssv0 + 0.5f * ssvT);
    const float jl = wsum(part_p) / p.fT + wsum(part_v) / p.fT;
    loss = fmaf(lam_jl, jl, fmaf(lam_sg, sg, toc));
    const bool ends_ok = (sqrtf(ssp0) < p.eps_pos) && (sqrtf(sspT) < p.eps_pos) &&
                         (sqrtf(ssv0) < p.eps_vel) && (sqrtf(ssvT) < p.eps_vel);
    ful = (ends_ok && __all_sync(FULL, lim_ok)) ? 1 : 0;
}

// ---------------------------------------------------------------------------
// Gradient phase: rows of  G_q = toc_g + lam_sg*sgp_g + lam_jl*jp_g  and
// G_v = lam_sg*sgv_g + lam_jl*jv_g  (trajectory.py:289-295, :91-126, robot.py:75-87),
// written as the operands of the backward contraction: XA = G_q, XB = -G_v
// (dK^T = -dK bit-exactly, checked in fgd_create()).
// ---------------------------------------------------------------------------
template <int RPL>
__device__ __forceinline__ void grad_phase(const DevParams &p, int lane, const Rows<RPL> &R, float lam_sg, float lam_jl,
                                           float4 *XA, float4 *XB)
{
    const int T = p.T;
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int t = lane + 32 * r;
        if (t >= T) continue;
        const float wt = (t == R.amax) ? (p.lam_max + p.w_avg) : p.w_avg;
        const float cgx = wt * R.gx[r], cgy = wt * R.gy[r];
        float xs[3], ys[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { xs[k] = -(p.link[k] * R.sn[r][k]); ys[k] = p.link[k] * R.cs[r][k]; }
        const float Sx = (xs[0] + xs[1]) + xs[2], Sy = (ys[0] + ys[1]) + ys[2];
        const float Cx[3] = {xs[0], xs[0] + xs[1], (xs[0] + xs[1]) + xs[2]};
        const float Cy[3] = {ys[0], ys[0] + ys[1], (ys[0] + ys[1]) + ys[2]};
        float gq[3], gv[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float Jx = (xs[k] + Sx) - Cx[k];
            const float Jy = (ys[k] + Sy) - Cy[k];
            const float tg = fmaf(cgy, Jy, cgx * Jx);
            const float qk = R.q[r][k], vk = R.v[r][k];
            const float sgp = (t == 0) ? R.d0[k] : ((t == T - 1) ? R.dT[k] : 0.0f);
            const float sgv = (t == 0 || t == T - 1) ? vk : 0.0f;
            const bool m = p.cvdl ? (qk > p.q_hi || qk < p.q_lo) : true;
            const float jpg = m ? ((qk - p.mean_q) * p.inv_std2) * p.inv_T : 0.0f;
            const bool mv = p.cvdl ? (fabsf(vk) > p.v_hi) : true;
            const float jvg = mv ? (vk * p.inv_vmax2) * p.inv_T : 0.0f;
            gq[k] = fmaf(lam_jl, jpg, fmaf(lam_sg, sgp, tg));
            gv[k] = fmaf(lam_jl, jvg, lam_sg * sgv);
        }
        XA[t] = make_float4(gq[0], gq[1], gq[2], 0.0f);
        XB[t] = make_float4(-gv[0], -gv[1], -gv[2], 0.0f);
    }
}

// alpha-gradient rows from the backward contraction: (K^T G_q + dK^T G_v) J^T
template <int RPL>
__device__ __forceinline__ void backward_rows(const DevParams &p, const float (&y1)[RPL][3], const float (&y2)[RPL][3],
                                              float (&g)[RPL][3])
{
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const float r0 = y1[r][0] + y2[r][0], r1 = y1[r][1] + y2[r][1], r2 = y1[r][2] + y2[r][2];
#pragma unroll
        for (int b = 0; b < 3; ++b) g[r][b] = fmaf(r2, p.J[b * 3 + 2], fmaf(r1, p.J[b * 3 + 1], r0 * p.J[b * 3]));
    }
}

}  // namespace fgd
