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

// ---------------------------------------------------------------------------
// Packed FP32 (sm_100 FFMA2 / FMUL2 / FADD2).  A lane's R adjacent time samples
// are handled as R/2 row pairs: .x = row 2p, .y = row 2p+1.  Every packed
// instruction is two independent IEEE-754 round-to-nearest operations, so the
// results are bit-identical to the scalar sequence of the mirror oracle while
// the instruction count per row halves.  Scalars (constants, obstacle
// coordinates, operand rows) enter through the .F32 broadcast operand form.
// ---------------------------------------------------------------------------
typedef float2 f2;
__device__ __forceinline__ f2 mk2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ f2 bc2(float a) { return make_float2(a, a); }
__device__ __forceinline__ f2 neg2(f2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return __fadd2_rn(a, neg2(b)); }      // a - b == a + (-b) bit for bit
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 sel2(bool mx, bool my, f2 a, f2 b) { return make_float2(mx ? a.x : b.x, my ? a.y : b.y); }
__device__ __forceinline__ f2 ss3_2(f2 a, f2 b, f2 c) { return fma2(c, c, fma2(b, b, mul2(a, a))); }

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

// sin and cos of two angles: Cody-Waite reduction by pi/2 (3 constants) and the
// cephes minimax polynomials on [-pi/4, pi/4]; <= 2 ulp for the angles a 3-joint
// arm with limits [-1, 2] rad produces.  Same operation sequence as the oracle.
__device__ __forceinline__ void quadrant(float j, float sn, float cs, float &S, float &C)
{
    const int n = ((int)j) & 3;
    S = (n & 1) ? cs : sn;
    C = (n & 1) ? sn : cs;
    if (n == 1 || n == 2) C = -C;
    if (n >= 2) S = -S;
}

__device__ __forceinline__ void sincos_cw2(f2 x, f2 &S, f2 &C)
{
    const f2 jj = mul2(x, bc2(6.366197467e-01f));
    const f2 j = mk2(rintf(jj.x), rintf(jj.y));
    f2 r = fma2(j, bc2(-1.570796371e+00f), x);
    r = fma2(j, bc2(4.371138829e-08f), r);
    r = fma2(j, bc2(1.715124510e-15f), r);
    const f2 s = mul2(r, r);
    f2 ps = fma2(s, bc2(-1.9515295891e-4f), bc2(8.3321608736e-3f));
    ps = fma2(ps, s, bc2(-1.6666654611e-1f));
    const f2 sn = fma2(mul2(ps, s), r, r);
    f2 pc = fma2(s, bc2(2.443315711809948e-5f), bc2(-1.388731625493765e-3f));
    pc = fma2(pc, s, bc2(4.166664568298827e-2f));
    const f2 cs = fma2(mul2(pc, s), s, fma2(bc2(-0.5f), s, bc2(1.0f)));
    quadrant(j.x, sn.x, cs.x, S.x, C.x);
    quadrant(j.y, sn.y, cs.y, S.y, C.y);
}

// ---------------------------------------------------------------------------
// RKHS contraction for the trajectory of this lane's group:
//   y1[r][a] = sum_k K [t_r][k] * x1[k][a]      (trajectory.py:65 / :295)
//   y2[r][a] = sum_k dK[t_r][k] * x2[k][a]
// k ascending, one fma per term.  KD is the interleaved operand table
//   KD[k][lane][0..R-1] = K[t_r][k],  KD[k][lane][R..2R-1] = dK[t_r][k]
// (row stride 2*TP floats, TP = LPT*R compile-time), so a lane fetches all its
// entries of column k with 2R/4 LDS.128 at immediate offsets and adjacent rows
// land in aligned register pairs: one FFMA2 per row pair and joint, the operand
// x[k][a] entering as a broadcast scalar.  The table comes from shared memory
// (KS) or, for T > 64, from L2 through the read-only path.
// SAME = true: x1 == x2 (forward evaluation), one operand load per k.
// ---------------------------------------------------------------------------
template <int LPT, int R, bool KS, bool SAME>
__device__ __forceinline__ void contract(const float *__restrict__ kd, int T,
                                         const float4 *__restrict__ x1, const float4 *__restrict__ x2,
                                         f2 (&y1)[R / 2][3], f2 (&y2)[R / 2][3])
{
    constexpr int RP = R / 2;
    constexpr int STRIDE = 2 * LPT * R;       // floats per column k
    constexpr int UNROLL = (R >= 8) ? 2 : ((R == 4) ? 3 : 5);
#pragma unroll
    for (int p = 0; p < RP; ++p)
#pragma unroll
        for (int a = 0; a < 3; ++a) { y1[p][a] = bc2(0.0f); y2[p][a] = bc2(0.0f); }

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
        for (int p = 0; p < RP; ++p) {
            const f2 kk = mk2(kv[2 * p], kv[2 * p + 1]), dk = mk2(kv[R + 2 * p], kv[R + 2 * p + 1]);
            y1[p][0] = fma2(kk, bc2(xa.x), y1[p][0]);
            y1[p][1] = fma2(kk, bc2(xa.y), y1[p][1]);
            y1[p][2] = fma2(kk, bc2(xa.z), y1[p][2]);
            y2[p][0] = fma2(dk, bc2(xb.x), y2[p][0]);
            y2[p][1] = fma2(dk, bc2(xb.y), y2[p][1]);
            y2[p][2] = fma2(dk, bc2(xb.z), y2[p][2]);
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
                                              const unsigned (&nz)[R], f2 (&y1)[R / 2][3], f2 (&y2)[R / 2][3])
{
    constexpr int RP = R / 2;
    constexpr int SO = LPT * R, SD = 2 * LPT * R;
    constexpr int UNROLL = (R >= 8) ? 2 : ((R == 4) ? 3 : 5);
#pragma unroll
    for (int p = 0; p < RP; ++p)
#pragma unroll
        for (int a = 0; a < 3; ++a) { y1[p][a] = bc2(0.0f); y2[p][a] = bc2(0.0f); }
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
        for (int p = 0; p < RP; ++p) {
            const f2 kk = mk2(kv[2 * p], kv[2 * p + 1]);
            y1[p][0] = fma2(kk, bc2(xa.x), y1[p][0]);
            y1[p][1] = fma2(kk, bc2(xa.y), y1[p][1]);
            y1[p][2] = fma2(kk, bc2(xa.z), y1[p][2]);
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
                const float2 *col = reinterpret_cast<const float2 *>(kd + (size_t)k * SD + R);
#pragma unroll
                for (int p = 0; p < RP; ++p) {
                    const f2 dv = KS ? col[p] : __ldg(col + p);
                    y2[p][0] = fma2(dv, bc2(xb.x), y2[p][0]);
                    y2[p][1] = fma2(dv, bc2(xb.y), y2[p][1]);
                    y2[p][2] = fma2(dv, bc2(xb.z), y2[p][2]);
                }
            }
        }
    }
}

// Per-lane row pairs kept between the cost phase and the gradient phase.  LEAN variants (R >= 4)
// keep only the obstacle gradient and recompute q, v and the sines/cosines in the gradient
// phase from the contraction rows (same arithmetic, same bits) to stay out of the spill zone.
template <int R, bool LEAN>
struct Rows {
    f2 q[LEAN ? 1 : R / 2][3], v[LEAN ? 1 : R / 2][3], sn[LEAN ? 1 : R / 2][3], cs[LEAN ? 1 : R / 2][3];
    f2 gx[R / 2], gy[R / 2];
    int amax;
};

// q, v of a row pair from the raw contraction rows ((M @ alpha) @ J, trajectory.py:65) and the
// sines / cosines of the cumulative joint angles (robot.py:32).
__device__ __forceinline__ void row_kinematics(const DevParams &p, const f2 (&yq)[3], const f2 (&yv)[3],
                                               f2 (&q)[3], f2 (&v)[3], f2 (&sn)[3], f2 (&cs)[3])
{
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        q[b] = fma2(yq[2], bc2(p.J[6 + b]), fma2(yq[1], bc2(p.J[3 + b]), mul2(yq[0], bc2(p.J[b]))));
        v[b] = fma2(yv[2], bc2(p.J[6 + b]), fma2(yv[1], bc2(p.J[3 + b]), mul2(yv[0], bc2(p.J[b]))));
    }
    const f2 c1 = q[0], c2 = add2(c1, q[1]), c3 = add2(c2, q[2]);
    sincos_cw2(c1, sn[0], cs[0]);
    sincos_cw2(c2, sn[1], cs[1]);
    sincos_cw2(c3, sn[2], cs[2]);
}

// U obstacles against RP row pairs (environment.py:32-58): sr += 1/den, sx += dx/den^2, sy += dy/den^2
// with den = 0.5 + 0.5 |f - o|^2.
template <int RP, int U, bool STRICT>
__device__ __forceinline__ void obstacle_block(const float2 *__restrict__ obs, const f2 (&x)[RP], const f2 (&y)[RP],
                                               f2 (&sr)[RP], f2 (&sx)[RP], f2 (&sy)[RP])
{
    float2 ob[U];
    if constexpr (U == 4) {
        const float4 a = *reinterpret_cast<const float4 *>(obs), b = *reinterpret_cast<const float4 *>(obs + 2);
        ob[0] = make_float2(a.x, a.y); ob[1] = make_float2(a.z, a.w); ob[2] = make_float2(b.x, b.y); ob[3] = make_float2(b.z, b.w);
    } else if constexpr (U == 2) {
        const float4 a = *reinterpret_cast<const float4 *>(obs);
        ob[0] = make_float2(a.x, a.y); ob[1] = make_float2(a.z, a.w);
    } else {
        ob[0] = obs[0];
    }
    f2 dx[U][RP], dy[U][RP], rr[U][RP];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int pr = 0; pr < RP; ++pr) {
            dx[u][pr] = add2(x[pr], bc2(-ob[u].x));
            dy[u][pr] = add2(y[pr], bc2(-ob[u].y));
            const f2 n = fma2(dy[u][pr], dy[u][pr], mul2(dx[u][pr], dx[u][pr]));
            rr[u][pr] = fma2(bc2(0.5f), n, bc2(0.5f));
        }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int pr = 0; pr < RP; ++pr) rr[u][pr] = mk2(rcp<STRICT>(rr[u][pr].x), rcp<STRICT>(rr[u][pr].y));
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int pr = 0; pr < RP; ++pr) {
            sr[pr] = add2(sr[pr], rr[u][pr]);
            const f2 r2 = mul2(rr[u][pr], rr[u][pr]);
            sx[pr] = fma2(r2, dx[u][pr], sx[pr]);
            sy[pr] = fma2(r2, dy[u][pr], sy[pr]);
        }
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
                                           const f2 (&yq)[R / 2][3], const f2 (&yv)[R / 2][3],
                                           const float *start, const float *goal, float lam_sg, float lam_jl,
                                           Rows<R, LEAN> &Rw, float &loss, float &toc, int &ful)
{
    constexpr int RP = R / 2;
    const int T = p.T;
    const int t0 = G.gl * R;
    const int lT = (T - 1) / R, rT = (T - 1) % R;
    float part_c = 0.0f, part_p = 0.0f, part_v = 0.0f, lmax = 0.0f;
    bool lim_ok = true;
    f2 x[RP], y[RP], sr[RP], sx[RP], sy[RP];
    float ssp0 = 0.0f, ssv0 = 0.0f, sspT = 0.0f, ssvT = 0.0f;      // meaningful in the lanes owning rows 0 / T-1
#pragma unroll
    for (int pr = 0; pr < RP; ++pr) {
        const bool valid0 = (t0 + 2 * pr) < T, valid1 = (t0 + 2 * pr + 1) < T;
        f2 q[3], v[3], sn[3], cs[3];
        row_kinematics(p, yq[pr], yv[pr], q, v, sn, cs);
        if constexpr (!LEAN) {
#pragma unroll
            for (int b = 0; b < 3; ++b) { Rw.q[pr][b] = q[b]; Rw.v[pr][b] = v[b]; Rw.sn[pr][b] = sn[b]; Rw.cs[pr][b] = cs[b]; }
        }
        x[pr] = fma2(bc2(p.link[2]), cs[2], fma2(bc2(p.link[1]), cs[1], mul2(bc2(p.link[0]), cs[0])));     // robot.py:33
        y[pr] = fma2(bc2(p.link[2]), sn[2], fma2(bc2(p.link[1]), sn[1], mul2(bc2(p.link[0]), sn[0])));     // robot.py:34
        sr[pr] = bc2(0.0f); sx[pr] = bc2(0.0f); sy[pr] = bc2(0.0f);
        // joint-limit penalties and limit predicates of these rows   trajectory.py:215-255, robot.py:104-113
        f2 e3[3], f3[3];
        bool ok0 = true, ok1 = true;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const f2 qb = q[b], vb = v[b];
            const f2 u = mul2(add2(qb, bc2(-p.mean_q)), bc2(p.inv_std));
            const f2 hu = mul2(bc2(0.5f), mul2(u, u));
            const bool m0 = p.cvdl ? (qb.x > p.q_hi || qb.x < p.q_lo) : true;
            const bool m1 = p.cvdl ? (qb.y > p.q_hi || qb.y < p.q_lo) : true;
            e3[b] = sel2(m0, m1, hu, bc2(0.0f));
            const f2 w = mul2(vb, bc2(p.inv_vmax));
            const f2 hw = mul2(bc2(0.5f), mul2(w, w));
            const bool n0 = p.cvdl ? (fabsf(vb.x) > p.v_hi) : true;
            const bool n1 = p.cvdl ? (fabsf(vb.y) > p.v_hi) : true;
            f3[b] = sel2(n0, n1, hw, bc2(0.0f));
            ok0 = ok0 & ((qb.x <= p.qmax) & (qb.x >= p.qmin) & (fabsf(vb.x) <= p.vmax));
            ok1 = ok1 & ((qb.y <= p.qmax) & (qb.y >= p.qmin) & (fabsf(vb.y) <= p.vmax));
        }
        lim_ok = lim_ok & (!valid0 | ok0) & (!valid1 | ok1);
        const f2 es = add2(add2(e3[0], e3[1]), e3[2]), fs = add2(add2(f3[0], f3[1]), f3[2]);
        if (valid0) { part_p = part_p + es.x; part_v = part_v + fs.x; }
        if (valid1) { part_p = part_p + es.y; part_v = part_v + fs.y; }
        // start / goal rows   trajectory.py:183-204
        if (pr == 0) {
            ssp0 = ss3(q[0].x - start[0], q[1].x - start[1], q[2].x - start[2]);
            ssv0 = ss3(v[0].x, v[1].x, v[2].x);
        }
        if (2 * pr == rT) {
            sspT = ss3(q[0].x - goal[0], q[1].x - goal[1], q[2].x - goal[2]);
            ssvT = ss3(v[0].x, v[1].x, v[2].x);
        }
        if (2 * pr + 1 == rT) {
            sspT = ss3(q[0].y - goal[0], q[1].y - goal[1], q[2].y - goal[2]);
            ssvT = ss3(v[0].y, v[1].y, v[2].y);
        }
    }
    // obstacle potential: all R samples of this lane against every obstacle, in blocks of U obstacles.
    // Each block runs in three stages (distances -> reciprocals -> accumulation) so that U * R/2
    // independent packed chains are in flight across the MUFU latency; the accumulation stage
    // visits the obstacles in ascending order (the oracle's summation order).
    const int n_obs = p.n_obs;
    constexpr int U = (R >= 8) ? 1 : ((R == 4) ? 2 : 4);
    int o = 0;
#pragma unroll 1
    for (; o + U <= n_obs; o += U) obstacle_block<RP, U, STRICT>(sObs + o, x, y, sr, sx, sy);
    if constexpr (U >= 4) {
        if (o + 2 <= n_obs) { obstacle_block<RP, 2, STRICT>(sObs + o, x, y, sr, sx, sy); o += 2; }
    }
    if constexpr (U >= 2) {
        if (o < n_obs) obstacle_block<RP, 1, STRICT>(sObs + o, x, y, sr, sx, sy);
    }
    f2 cost[RP];
#pragma unroll
    for (int pr = 0; pr < RP; ++pr) {
        const f2 c = mul2(bc2(0.8f), sr[pr]);
        cost[pr] = c; Rw.gx[pr] = mul2(bc2(-0.8f), sx[pr]); Rw.gy[pr] = mul2(bc2(-0.8f), sy[pr]);
        if ((t0 + 2 * pr) < T) {
            part_c = part_c + c.x;
            lmax = fmaxf(lmax, c.x);          // c >= 0
        }
        if ((t0 + 2 * pr + 1) < T) {
            part_c = part_c + c.y;
            lmax = fmaxf(lmax, c.y);
        }
    }
    // max / first argmax / mean over t
    const float maxc = gmax<LPT>(lmax);
    int cand = 0x7fffffff;
#pragma unroll
    for (int pr = RP - 1; pr >= 0; --pr) {
        if ((t0 + 2 * pr + 1) < T && cost[pr].y == maxc) cand = t0 + 2 * pr + 1;
        if ((t0 + 2 * pr) < T && cost[pr].x == maxc) cand = t0 + 2 * pr;
    }
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
                                           const f2 (&yq)[R / 2][3], const f2 (&yv)[R / 2][3], const float *start, const float *goal,
                                           float lam_sg, float lam_jl, float4 *XA, float4 *XB, bool commit, unsigned (&nz)[R])
{
    constexpr int RP = R / 2;
    const int T = p.T;
#pragma unroll
    for (int pr = 0; pr < RP; ++pr) {
        const int ta = G.gl * R + 2 * pr, tb = ta + 1;
        f2 q[3], v[3], sn[3], cs[3];
        if constexpr (LEAN) {
            row_kinematics(p, yq[pr], yv[pr], q, v, sn, cs);
        } else {
#pragma unroll
            for (int b = 0; b < 3; ++b) { q[b] = Rw.q[pr][b]; v[b] = Rw.v[pr][b]; sn[b] = Rw.sn[pr][b]; cs[b] = Rw.cs[pr][b]; }
        }
        const float w_hi = p.lam_max + p.w_avg;
        const f2 wt = mk2((ta == Rw.amax) ? w_hi : p.w_avg, (tb == Rw.amax) ? w_hi : p.w_avg);
        const f2 cgx = mul2(wt, Rw.gx[pr]), cgy = mul2(wt, Rw.gy[pr]);
        f2 xs[3], ys[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { xs[k] = neg2(mul2(bc2(p.link[k]), sn[k])); ys[k] = mul2(bc2(p.link[k]), cs[k]); }
        const f2 Sx = add2(add2(xs[0], xs[1]), xs[2]), Sy = add2(add2(ys[0], ys[1]), ys[2]);
        const f2 Cx[3] = {xs[0], add2(xs[0], xs[1]), add2(add2(xs[0], xs[1]), xs[2])};
        const f2 Cy[3] = {ys[0], add2(ys[0], ys[1]), add2(add2(ys[0], ys[1]), ys[2])};
        f2 gq[3], gv[3];
        const bool a0 = (ta == 0), aT = (ta == T - 1), b0 = false, bT = (tb == T - 1);     // tb >= 1
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const f2 Jx = sub2(add2(xs[k], Sx), Cx[k]);
            const f2 Jy = sub2(add2(ys[k], Sy), Cy[k]);
            const f2 tg = fma2(cgy, Jy, mul2(cgx, Jx));
            const f2 qk = q[k], vk = v[k];
            f2 sgp, sgv;
            sgp.x = a0 ? (qk.x - start[k]) : (aT ? (qk.x - goal[k]) : 0.0f);
            sgp.y = b0 ? (qk.y - start[k]) : (bT ? (qk.y - goal[k]) : 0.0f);
            sgv.x = (a0 || aT) ? vk.x : 0.0f;
            sgv.y = (b0 || bT) ? vk.y : 0.0f;
            const bool m0 = p.cvdl ? (qk.x > p.q_hi || qk.x < p.q_lo) : true;
            const bool m1 = p.cvdl ? (qk.y > p.q_hi || qk.y < p.q_lo) : true;
            const f2 jpg = sel2(m0, m1, mul2(mul2(add2(qk, bc2(-p.mean_q)), bc2(p.inv_std2)), bc2(p.inv_T)), bc2(0.0f));
            const bool n0 = p.cvdl ? (fabsf(vk.x) > p.v_hi) : true;
            const bool n1 = p.cvdl ? (fabsf(vk.y) > p.v_hi) : true;
            const f2 jvg = sel2(n0, n1, mul2(mul2(vk, bc2(p.inv_vmax2)), bc2(p.inv_T)), bc2(0.0f));
            gq[k] = fma2(bc2(lam_jl), jpg, fma2(bc2(lam_sg), sgp, tg));
            gv[k] = fma2(bc2(lam_jl), jvg, mul2(bc2(lam_sg), sgv));
        }
        const bool wa = commit && ta < T, wb = commit && tb < T;
        if (wa) {
            XA[ta] = make_float4(gq[0].x, gq[1].x, gq[2].x, 0.0f);
            XB[ta] = make_float4(-gv[0].x, -gv[1].x, -gv[2].x, 0.0f);
        }
        if (wb) {
            XA[tb] = make_float4(gq[0].y, gq[1].y, gq[2].y, 0.0f);
            XB[tb] = make_float4(-gv[0].y, -gv[1].y, -gv[2].y, 0.0f);
        }
        const unsigned ma = __ballot_sync(FULL, wa && (gv[0].x != 0.0f || gv[1].x != 0.0f || gv[2].x != 0.0f));
        const unsigned mb = __ballot_sync(FULL, wb && (gv[0].y != 0.0f || gv[1].y != 0.0f || gv[2].y != 0.0f));
        if (commit) { nz[2 * pr] = ma; nz[2 * pr + 1] = mb; }
    }
}

// alpha-gradient rows from the backward contraction: (K^T G_q + dK^T G_v) J^T
template <int R>
__device__ __forceinline__ void backward_rows(const DevParams &p, const f2 (&y1)[R / 2][3], const f2 (&y2)[R / 2][3],
                                              f2 (&g)[R / 2][3])
{
#pragma unroll
    for (int pr = 0; pr < R / 2; ++pr) {
        const f2 r0 = add2(y1[pr][0], y2[pr][0]), r1 = add2(y1[pr][1], y2[pr][1]), r2 = add2(y1[pr][2], y2[pr][2]);
#pragma unroll
        for (int b = 0; b < 3; ++b)
            g[pr][b] = fma2(r2, bc2(p.J[b * 3 + 2]), fma2(r1, bc2(p.J[b * 3 + 1]), mul2(r0, bc2(p.J[b * 3]))));
    }
}

}  // namespace fgd
