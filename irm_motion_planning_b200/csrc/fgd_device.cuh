// fgd_device.cuh -- device-side building blocks of the batched FGD iteration (sm_100a).
//
// Mapping.  A trajectory is owned by a GROUP of LPT lanes.
// The shipped configurations use LPT = 32: one warp per trajectory.  (The code is
// written for LPT = 8/16 as well -- 32/LPT trajectories side by side in a warp --
// but on B200 those variants deadlocked at full-mask collectives after intra-warp
// divergence and are not instantiated; see DESIGN.md "open issues".)
// Lane l of the group owns the R adjacent time samples t = R*l .. R*l+R-1
// (R = 2, 4 or 8).  The RKHS contraction produces rows t of q = K alpha J and
// v = dK alpha J in the lane that then does forward kinematics, the obstacle
// potential and the penalty terms for those samples -- no shared-memory round
// trip between the two.  K and dK are staged once per CTA in shared memory,
// transposed, so a lane fetches its R row entries of column k with one or two
// LDS.128 and the groups of a warp share the load by broadcast; the operand rows
// (alpha' or the q/v-gradients) sit in per-group shared buffers and are broadcast
// inside the group.  Reductions over t: lane-serial over the R rows, then an
// xor butterfly over the LPT lanes -- the order the mirror oracle reproduces.
//
// The groups of a warp execute the contraction, the cost phase, the gradient
// phase and the normalisation as ONE converged instruction stream; group-specific
// decisions only predicate what is committed.
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
    const float *KD;         // [T][LPT][2R]: K and dK entries of column k interleaved per lane (zero padded rows >= T)
    const float *KO;         // [T][LPT][R]:  K entries only (dense half of the backward contraction)
    const float *obs;        // [n_obs][2]
    float *alpha;            // [B][T][3]
    const float *start, *goal;
    float *fstate;
    int *istate;
    unsigned *queue;
    int *dbg;                // optional host-mapped progress markers (debug builds only)
};

struct EvalPtrs {
    float lam_sg, lam_jl;
    float *loss, *toc, *grad, *q, *v;
    int *fulfilled;
};

// group-uniform per-trajectory scalars (shared memory, one per group)
struct Slot {
    int traj, status, outer, inner, inner_total, cand_evals, accepts, ful, j, done_iters;
    unsigned hash;
    float lam_sg, lam_jl, lr, loss, toc, alpha_norm, last_new;
    float start[3], goal[3];
};

// lane geometry of one trajectory group
template <int LPT>
struct Group {
    int lane, gl, base;
    unsigned mask;
    __device__ __forceinline__ Group()
    {
        lane = threadIdx.x & 31;
        gl = lane & (LPT - 1);
        base = lane & ~(LPT - 1);
        mask = (LPT == 32) ? FULL : (((1u << LPT) - 1u) << base);
    }
};

// Group collectives.  They are always executed by all 32 lanes of the warp (the kernels keep
// the groups of a warp converged around them), so the full mask with width = LPT is legal and
// compiles to one SHFL per step.
template <int LPT>
__device__ __forceinline__ float gsum(float v)
{
#pragma unroll
    for (int o = LPT / 2; o >= 1; o >>= 1) v = v + __shfl_xor_sync(FULL, v, o, LPT);
    return v;
}

template <int LPT>
__device__ __forceinline__ float gmax(float v)
{
#pragma unroll
    for (int o = LPT / 2; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o, LPT));
    return v;
}

template <int LPT>
__device__ __forceinline__ int gmin_int(int v)
{
#pragma unroll
    for (int o = LPT / 2; o >= 1; o >>= 1) v = min(v, __shfl_xor_sync(FULL, v, o, LPT));
    return v;
}

template <int LPT>
__device__ __forceinline__ float gbcast(float v, int src_lane_in_group) { return __shfl_sync(FULL, v, src_lane_in_group, LPT); }

template <int LPT>
__device__ __forceinline__ bool gall(bool pred, const Group<LPT> &G)
{
    const unsigned b = __ballot_sync(FULL, pred);
    return (b & G.mask) == G.mask;
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
// RKHS contraction for the trajectory of this lane's group:
//   y1[r][a] = sum_k K [t_r][k] * x1[k][a]      (trajectory.py:65 / :295)
//   y2[r][a] = sum_k dK[t_r][k] * x2[k][a]
// k ascending, one fmaf per term.  KD is the interleaved operand table
//   KD[k][lane][0..R-1] = K[t_r][k],  KD[k][lane][R..2R-1] = dK[t_r][k]
// (row stride 2*TP floats, TP = LPT*R compile-time), so a lane fetches all its
// entries of column k with 2R/4 LDS.128 at immediate offsets.  It comes from
// shared memory (KS) or, for T > 128, from L2 through the read-only path.
// SAME = true: x1 == x2 (forward evaluation), one operand load per k.
// ---------------------------------------------------------------------------
template <int LPT, int R, bool KS, bool SAME>
__device__ __forceinline__ void contract(const float *__restrict__ kd, int T,
                                         const float4 *__restrict__ x1, const float4 *__restrict__ x2,
                                         float (&y1)[R][3], float (&y2)[R][3])
{
    constexpr int STRIDE = 2 * LPT * R;       // floats per column k
    constexpr int UNROLL = (R >= 8) ? 2 : ((R == 4) ? 3 : 5);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int a = 0; a < 3; ++a) { y1[r][a] = 0.0f; y2[r][a] = 0.0f; }

#pragma unroll UNROLL
    for (int k = 0; k < T; ++k) {
        float kv[2 * R];
#pragma unroll
        for (int c = 0; c < 2 * R; c += 4) {
            float4 v;
            if constexpr (KS) v = *reinterpret_cast<const float4 *>(kd + (size_t)k * STRIDE + c);
            else v = __ldg(reinterpret_cast<const float4 *>(kd + (size_t)k * STRIDE + c));
            kv[c] = v.x; kv[c + 1] = v.y; kv[c + 2] = v.z; kv[c + 3] = v.w;
        }
        const float4 xa = x1[k];
        float4 xb = xa;
        if constexpr (!SAME) xb = x2[k];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            y1[r][0] = fmaf(kv[r], xa.x, y1[r][0]);
            y1[r][1] = fmaf(kv[r], xa.y, y1[r][1]);
            y1[r][2] = fmaf(kv[r], xa.z, y1[r][2]);
            y2[r][0] = fmaf(kv[R + r], xb.x, y2[r][0]);
            y2[r][1] = fmaf(kv[R + r], xb.y, y2[r][1]);
            y2[r][2] = fmaf(kv[R + r], xb.z, y2[r][2]);
        }
    }
}

// ---------------------------------------------------------------------------
// Backward contraction  y1 = K G_q (dense),  y2 = dK (-G_v) (sparse).
// G_v = lam_sg*sgv_g + lam_jl*jv_g is zero except in rows 0 and T-1 and where the
// velocity limit is violated (trajectory.py:207-212, 258-268), so the dK half only
// visits the rows flagged in nz[] (bit l of nz[r] <=> row R*l + r is non-zero).
// Skipped terms are exact zeros, so the result equals the dense sum bit for bit;
// the visited terms are still accumulated in ascending k.
// ---------------------------------------------------------------------------
template <int LPT, int R, bool KS>
__device__ __forceinline__ void contract_back(const float *__restrict__ ko, const float *__restrict__ kd, int T,
                                              const float4 *__restrict__ xa_rows, const float4 *__restrict__ xb_rows,
                                              const unsigned (&nz)[R], float (&y1)[R][3], float (&y2)[R][3])
{
    constexpr int SO = LPT * R, SD = 2 * LPT * R;
    constexpr int UNROLL = (R >= 8) ? 2 : ((R == 4) ? 3 : 5);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int a = 0; a < 3; ++a) { y1[r][a] = 0.0f; y2[r][a] = 0.0f; }
#pragma unroll UNROLL
    for (int k = 0; k < T; ++k) {
        float kv[R];
        if constexpr (R == 2) {
            float2 v;
            if constexpr (KS) v = *reinterpret_cast<const float2 *>(ko + (size_t)k * SO);
            else v = __ldg(reinterpret_cast<const float2 *>(ko + (size_t)k * SO));
            kv[0] = v.x; kv[1] = v.y;
        } else {
#pragma unroll
            for (int c = 0; c < R; c += 4) {
                float4 v;
                if constexpr (KS) v = *reinterpret_cast<const float4 *>(ko + (size_t)k * SO + c);
                else v = __ldg(reinterpret_cast<const float4 *>(ko + (size_t)k * SO + c));
                kv[c] = v.x; kv[c + 1] = v.y; kv[c + 2] = v.z; kv[c + 3] = v.w;
            }
        }
        const float4 xa = xa_rows[k];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            y1[r][0] = fmaf(kv[r], xa.x, y1[r][0]);
            y1[r][1] = fmaf(kv[r], xa.y, y1[r][1]);
            y1[r][2] = fmaf(kv[r], xa.z, y1[r][2]);
        }
    }
    unsigned any = 0u;
#pragma unroll
    for (int r = 0; r < R; ++r) any |= nz[r];
    while (any) {                                   // warp-uniform: ascending lane, then ascending r = ascending k
        const int l = __ffs(any) - 1;
        any &= any - 1;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if ((nz[r] >> l) & 1u) {
                const int k = l * R + r;
                const float4 xb = xb_rows[k];
                const float *col = kd + (size_t)k * SD + R;
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const float dv = KS ? col[q] : __ldg(col + q);
                    y2[q][0] = fmaf(dv, xb.x, y2[q][0]);
                    y2[q][1] = fmaf(dv, xb.y, y2[q][1]);
                    y2[q][2] = fmaf(dv, xb.z, y2[q][2]);
                }
            }
        }
    }
}

// Per-lane rows kept between the cost phase and the gradient phase.  LEAN variants (R >= 4)
// keep only the obstacle gradient and recompute q, v and the sines/cosines in the gradient
// phase from the contraction rows (same arithmetic, same bits) to stay out of the spill zone.
template <int R, bool LEAN>
struct Rows {
    float q[LEAN ? 1 : R][3], v[LEAN ? 1 : R][3], sn[LEAN ? 1 : R][3], cs[LEAN ? 1 : R][3];
    float gx[R], gy[R];
    int amax;
};

// q, v rows from the raw contraction rows ((M @ alpha) @ J, trajectory.py:65) and the
// sines / cosines of the cumulative joint angles (robot.py:32).
__device__ __forceinline__ void row_kinematics(const DevParams &p, const float (&yq)[3], const float (&yv)[3],
                                               float (&q)[3], float (&v)[3], float (&sn)[3], float (&cs)[3])
{
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        q[b] = fmaf(yq[2], p.J[6 + b], fmaf(yq[1], p.J[3 + b], yq[0] * p.J[b]));
        v[b] = fmaf(yv[2], p.J[6 + b], fmaf(yv[1], p.J[3 + b], yv[0] * p.J[b]));
    }
    const float c1 = q[0], c2 = c1 + q[1], c3 = c2 + q[2];
    sincos_cw(c1, sn[0], cs[0]);
    sincos_cw(c2, sn[1], cs[1]);
    sincos_cw(c3, sn[2], cs[2]);
}

// ---------------------------------------------------------------------------
// Cost phase: compute_trajectory_cost + constraintsFulfilled for one trajectory
// whose raw contraction rows are yq (K alpha) and yv (dK alpha).
//   trajectory.py:271-281 (total), :81-88 (max/mean), :183-255 (penalties),
//   :129-137 + robot.py:90-113 (constraint predicates), robot.py:29-36 (fk),
//   environment.py:32-58 (obstacle potential and its (x,y)-gradient).
// The obstacle loop accumulates sum 1/den and sum d/den^2; the constant factors
// 0.8 and -0.8 of environment.py:43,57 are applied once per sample.
// ---------------------------------------------------------------------------
template <int LPT, int R, bool STRICT, bool LEAN>
__device__ __forceinline__ void cost_phase(const DevParams &p, const float2 *__restrict__ sObs, const Group<LPT> &G,
                                           const float (&yq)[R][3], const float (&yv)[R][3],
                                           const float *start, const float *goal, float lam_sg, float lam_jl,
                                           Rows<R, LEAN> &Rw, float &loss, float &toc, int &ful)
{
    const int T = p.T;
    const int t0 = G.gl * R;
    const int lT = (T - 1) / R, rT = (T - 1) % R;
    float part_c = 0.0f, part_p = 0.0f, part_v = 0.0f, lmax = 0.0f;
    bool lim_ok = true;
    float x[R], y[R], sr[R], sx[R], sy[R];
    float ssp0 = 0.0f, ssv0 = 0.0f, sspT = 0.0f, ssvT = 0.0f;      // meaningful in the lanes owning rows 0 / T-1
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool valid = (t0 + r) < T;
        float q[3], v[3], sn[3], cs[3];
        row_kinematics(p, yq[r], yv[r], q, v, sn, cs);
        if constexpr (!LEAN) {
#pragma unroll
            for (int b = 0; b < 3; ++b) { Rw.q[r][b] = q[b]; Rw.v[r][b] = v[b]; Rw.sn[r][b] = sn[b]; Rw.cs[r][b] = cs[b]; }
        }
        x[r] = fmaf(p.link[2], cs[2], fmaf(p.link[1], cs[1], p.link[0] * cs[0]));     // robot.py:33
        y[r] = fmaf(p.link[2], sn[2], fmaf(p.link[1], sn[1], p.link[0] * sn[0]));     // robot.py:34
        sr[r] = 0.0f; sx[r] = 0.0f; sy[r] = 0.0f;
        // joint-limit penalties and limit predicates of this row   trajectory.py:215-255, robot.py:104-113
        float e3[3], f3[3];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const float qb = q[b], vb = v[b];
            const float u = (qb - p.mean_q) * p.inv_std;
            const bool m = p.cvdl ? (qb > p.q_hi || qb < p.q_lo) : true;
            e3[b] = m ? 0.5f * (u * u) : 0.0f;
            const float w = vb * p.inv_vmax;
            const bool mv = p.cvdl ? (fabsf(vb) > p.v_hi) : true;
            f3[b] = mv ? 0.5f * (w * w) : 0.0f;
            lim_ok = lim_ok & (!valid | ((qb <= p.qmax) & (qb >= p.qmin) & (fabsf(vb) <= p.vmax)));
        }
        if (valid) {
            part_p = part_p + ((e3[0] + e3[1]) + e3[2]);
            part_v = part_v + ((f3[0] + f3[1]) + f3[2]);
        }
        // start / goal rows   trajectory.py:183-204
        if (r == 0) {
            ssp0 = ss3(q[0] - start[0], q[1] - start[1], q[2] - start[2]);
            ssv0 = ss3(v[0], v[1], v[2]);
        }
        if (r == rT) {
            sspT = ss3(q[0] - goal[0], q[1] - goal[1], q[2] - goal[2]);
            ssvT = ss3(v[0], v[1], v[2]);
        }
    }
    // obstacle potential: all R samples of this lane against every obstacle
    const int n_obs = p.n_obs;
    constexpr int OBS_UNROLL = (R >= 8) ? 1 : 2;      // R independent chains per obstacle already
    float2 ob_next = sObs[0];                       // software pipelining: the next obstacle is fetched one trip ahead
#pragma unroll OBS_UNROLL
    for (int o = 0; o < n_obs; ++o) {
        const float2 ob = ob_next;
        ob_next = sObs[o + 1];                      // the buffer is padded by one pair (make_layout)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float dx = x[r] - ob.x, dy = y[r] - ob.y;
            const float n = fmaf(dy, dy, dx * dx);
            const float den = fmaf(0.5f, n, 0.5f);
            const float rr = rcp<STRICT>(den);
            sr[r] = sr[r] + rr;
            const float r2 = rr * rr;
            sx[r] = fmaf(r2, dx, sx[r]);
            sy[r] = fmaf(r2, dy, sy[r]);
        }
    }
    float cost[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float c = 0.8f * sr[r];
        cost[r] = c; Rw.gx[r] = -0.8f * sx[r]; Rw.gy[r] = -0.8f * sy[r];
        if ((t0 + r) < T) {
            part_c = part_c + c;
            lmax = fmaxf(lmax, c);          // c >= 0
        }
    }
    // max / first argmax / mean over t
    const float maxc = gmax<LPT>(lmax);
    int cand = 0x7fffffff;
#pragma unroll
    for (int r = R - 1; r >= 0; --r)
        if ((t0 + r) < T && cost[r] == maxc) cand = t0 + r;
    Rw.amax = gmin_int<LPT>(cand);
    const float avg = gsum<LPT>(part_c) / p.fT;
    toc = fmaf(p.lam_max, maxc, p.oml * avg);

    ssp0 = gbcast<LPT>(ssp0, 0);
    ssv0 = gbcast<LPT>(ssv0, 0);
    sspT = gbcast<LPT>(sspT, lT);
    ssvT = gbcast<LPT>(ssvT, lT);
    const float sg = (0.5f * ssp0 + 0.5f * sspT) + (0.5f * ssv0 + 0.5f * ssvT);
    const float jl = gsum<LPT>(part_p) / p.fT + gsum<LPT>(part_v) / p.fT;
    loss = fmaf(lam_jl, jl, fmaf(lam_sg, sg, toc));
    const bool ends_ok = (sqrtf(ssp0) < p.eps_pos) && (sqrtf(sspT) < p.eps_pos) &&
                         (sqrtf(ssv0) < p.eps_vel) && (sqrtf(ssvT) < p.eps_vel);
    ful = (ends_ok && gall<LPT>(lim_ok, G)) ? 1 : 0;
}

// ---------------------------------------------------------------------------
// Gradient phase: rows of  G_q = toc_g + lam_sg*sgp_g + lam_jl*jp_g  and
// G_v = lam_sg*sgv_g + lam_jl*jv_g  (trajectory.py:289-295, :91-126, robot.py:75-87),
// written as the operands of the backward contraction: XA = G_q, XB = -G_v
// (dK^T = -dK bit-exactly, checked in fgd_create()).  nz[r] collects the rows whose
// velocity gradient is not identically zero (consumed by contract_back).
// ---------------------------------------------------------------------------
template <int LPT, int R, bool LEAN>
__device__ __forceinline__ void grad_phase(const DevParams &p, const Group<LPT> &G, const Rows<R, LEAN> &Rw,
                                           const float (&yq)[R][3], const float (&yv)[R][3], const float *start, const float *goal,
                                           float lam_sg, float lam_jl, float4 *XA, float4 *XB, bool commit, unsigned (&nz)[R])
{
    const int T = p.T;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int t = G.gl * R + r;
        float q[3], v[3], sn[3], cs[3];
        if constexpr (LEAN) {
            row_kinematics(p, yq[r], yv[r], q, v, sn, cs);
        } else {
#pragma unroll
            for (int b = 0; b < 3; ++b) { q[b] = Rw.q[r][b]; v[b] = Rw.v[r][b]; sn[b] = Rw.sn[r][b]; cs[b] = Rw.cs[r][b]; }
        }
        const float wt = (t == Rw.amax) ? (p.lam_max + p.w_avg) : p.w_avg;
        const float cgx = wt * Rw.gx[r], cgy = wt * Rw.gy[r];
        float xs[3], ys[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { xs[k] = -(p.link[k] * sn[k]); ys[k] = p.link[k] * cs[k]; }
        const float Sx = (xs[0] + xs[1]) + xs[2], Sy = (ys[0] + ys[1]) + ys[2];
        const float Cx[3] = {xs[0], xs[0] + xs[1], (xs[0] + xs[1]) + xs[2]};
        const float Cy[3] = {ys[0], ys[0] + ys[1], (ys[0] + ys[1]) + ys[2]};
        float gq[3], gv[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float Jx = (xs[k] + Sx) - Cx[k];
            const float Jy = (ys[k] + Sy) - Cy[k];
            const float tg = fmaf(cgy, Jy, cgx * Jx);
            const float qk = q[k], vk = v[k];
            const float sgp = (t == 0) ? (qk - start[k]) : ((t == T - 1) ? (qk - goal[k]) : 0.0f);
            const float sgv = (t == 0 || t == T - 1) ? vk : 0.0f;
            const bool m = p.cvdl ? (qk > p.q_hi || qk < p.q_lo) : true;
            const float jpg = m ? ((qk - p.mean_q) * p.inv_std2) * p.inv_T : 0.0f;
            const bool mv = p.cvdl ? (fabsf(vk) > p.v_hi) : true;
            const float jvg = mv ? (vk * p.inv_vmax2) * p.inv_T : 0.0f;
            gq[k] = fmaf(lam_jl, jpg, fmaf(lam_sg, sgp, tg));
            gv[k] = fmaf(lam_jl, jvg, lam_sg * sgv);
        }
        const bool write = commit && t < T;
        if (write) {
            XA[t] = make_float4(gq[0], gq[1], gq[2], 0.0f);
            XB[t] = make_float4(-gv[0], -gv[1], -gv[2], 0.0f);
        }
        const unsigned m = __ballot_sync(FULL, write && (gv[0] != 0.0f || gv[1] != 0.0f || gv[2] != 0.0f));
        if (commit) nz[r] = m;
    }
}

// alpha-gradient rows from the backward contraction: (K^T G_q + dK^T G_v) J^T
template <int R>
__device__ __forceinline__ void backward_rows(const DevParams &p, const float (&y1)[R][3], const float (&y2)[R][3],
                                              float (&g)[R][3])
{
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float r0 = y1[r][0] + y2[r][0], r1 = y1[r][1] + y2[r][1], r2 = y1[r][2] + y2[r][2];
#pragma unroll
        for (int b = 0; b < 3; ++b) g[r][b] = fmaf(r2, p.J[b * 3 + 2], fmaf(r1, p.J[b * 3 + 1], r0 * p.J[b * 3]));
    }
}

}  // namespace fgd
