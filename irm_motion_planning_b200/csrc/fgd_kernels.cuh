// fgd_kernels.cuh -- the persistent optimiser kernel, the evaluation kernel and
// the restart-argmin kernel.  See fgd_device.cuh for the mapping.
#pragma once
#include "fgd_device.cuh"
#include "../../include/fgd_b200.h"

namespace fgd {

// shared-memory carve-up, identical on host and device
struct SmemLayout {
    int k_floats;     // floats of staged K (and dK), 0 when K stays in L2
    int obs_pairs;    // obstacle slots (padded to even)
    int x_rows;       // float4 rows per operand buffer (= T)
    int n_slots;      // NW * S
    __host__ __device__ size_t bytes() const
    {
        return (size_t)2 * k_floats * 4 + (size_t)obs_pairs * 8 + (size_t)n_slots * 2 * x_rows * 16 + (size_t)n_slots * sizeof(Slot);
    }
};

__host__ __device__ inline SmemLayout make_layout(int T, int TP, int n_obs, bool ks, int n_slots)
{
    SmemLayout l;
    l.k_floats = ks ? T * TP : 0;
    l.obs_pairs = (n_obs + 1) & ~1;
    l.x_rows = T;
    l.n_slots = n_slots;
    return l;
}

__device__ __forceinline__ void hash_step(Slot &st, unsigned code) { st.hash = st.hash * 1000003u + code; }

// candidate  (1 - lam_reg*lr) * alpha - lr * dir      optimizer_BLS.py:139, optimizer_GD.py:185
template <int RPL>
__device__ __forceinline__ void write_candidate(const DevParams &p, int lane, float lr, const float (&a)[RPL][3],
                                                const float (&d)[RPL][3], float4 *XA)
{
    const float c1 = 1.0f - p.lam_reg * lr;
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int t = lane + 32 * r;
        if (t < p.T)
            XA[t] = make_float4(fmaf(c1, a[r][0], -(lr * d[r][0])), fmaf(c1, a[r][1], -(lr * d[r][1])),
                                fmaf(c1, a[r][2], -(lr * d[r][2])), 0.0f);
    }
}

template <int RPL>
__device__ __forceinline__ void accept_candidate(const DevParams &p, float lr, float (&a)[RPL][3], const float (&d)[RPL][3])
{
    const float c1 = 1.0f - p.lam_reg * lr;
#pragma unroll
    for (int r = 0; r < RPL; ++r)
#pragma unroll
        for (int b = 0; b < 3; ++b) a[r][b] = fmaf(c1, a[r][b], -(lr * d[r][b]));
}

template <int RPL>
__device__ __forceinline__ void save_slot(const DevParams &p, int lane, const Slot &st, int status, const float (&a)[RPL][3])
{
    const int b = st.traj;
    float *ap = p.alpha + (size_t)b * p.T * 3;
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int t = lane + 32 * r;
        if (t < p.T) { ap[t * 3] = a[r][0]; ap[t * 3 + 1] = a[r][1]; ap[t * 3 + 2] = a[r][2]; }
    }
    if (lane == 0) {
        float *fs = p.fstate + (size_t)b * FGD_FSTATE;
        int *is = p.istate + (size_t)b * FGD_ISTATE;
        fs[FGD_F_LAM_SG] = st.lam_sg; fs[FGD_F_LAM_JL] = st.lam_jl; fs[FGD_F_LR] = st.lr;
        fs[FGD_F_LOSS] = st.loss; fs[FGD_F_TOC] = st.toc; fs[FGD_F_LAST_NEW_LOSS] = st.last_new;
        is[FGD_I_STATUS] = status; is[FGD_I_OUTER] = st.outer; is[FGD_I_INNER] = st.inner;
        is[FGD_I_INNER_TOTAL] = st.inner_total; is[FGD_I_CAND_EVALS] = st.cand_evals; is[FGD_I_ACCEPTS] = st.accepts;
        is[FGD_I_FULFILLED] = st.ful; is[FGD_I_HASH] = (int)st.hash;
    }
}

// Start (or restart after a lambda increase / a resumed launch) with the loss and
// gradient operands at the current alpha: optimizer_BLS.py:163, optimizer_GD.py:210.
template <int RPL>
__device__ __forceinline__ void begin_outer_eval(const DevParams &p, int lane, Slot &st, int &kind, const float (&a)[RPL][3], float4 *XA)
{
    if (p.mode == 1) st.lr = p.gd_lr[st.outer];
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int t = lane + 32 * r;
        if (t < p.T) XA[t] = make_float4(a[r][0], a[r][1], a[r][2], 0.0f);
    }
    kind = K_EVAL0;
}

// Pull the next unfinished trajectory from the batch queue into this slot.
template <int RPL>
__device__ __forceinline__ void fetch_slot(const DevParams &p, int lane, Slot &st, int &kind, float (&a)[RPL][3], float4 *XA)
{
    for (;;) {
        unsigned idx = 0;
        if (lane == 0) idx = atomicAdd(p.queue, 1u);
        idx = __shfl_sync(FULL, idx, 0);
        if (idx >= (unsigned)p.B) { st.traj = -1; kind = K_IDLE; return; }
        const int *is = p.istate + (size_t)idx * FGD_ISTATE;
        const int status = is[FGD_I_STATUS];
        if (status == FGD_ST_DONE) continue;
        const float *fs = p.fstate + (size_t)idx * FGD_FSTATE;
        st.traj = (int)idx;
        st.done_iters = 0; st.j = 0; st.alpha_norm = 0.0f;
        if (status == FGD_ST_FRESH) {
            st.lam_sg = p.lam_sg0; st.lam_jl = p.lam_jl0;
            st.lr = (p.mode == 0) ? p.bls_lr0 : p.gd_lr[0];
            st.outer = 0; st.inner = 0; st.inner_total = 0; st.cand_evals = 0; st.accepts = 0; st.ful = 0; st.hash = 0u;
            st.loss = 0.0f; st.toc = 0.0f; st.last_new = 0.0f;
        } else {
            st.lam_sg = fs[FGD_F_LAM_SG]; st.lam_jl = fs[FGD_F_LAM_JL]; st.lr = fs[FGD_F_LR];
            st.loss = fs[FGD_F_LOSS]; st.toc = fs[FGD_F_TOC]; st.last_new = fs[FGD_F_LAST_NEW_LOSS];
            st.outer = is[FGD_I_OUTER]; st.inner = is[FGD_I_INNER]; st.inner_total = is[FGD_I_INNER_TOTAL];
            st.cand_evals = is[FGD_I_CAND_EVALS]; st.accepts = is[FGD_I_ACCEPTS]; st.ful = is[FGD_I_FULFILLED];
            st.hash = (unsigned)is[FGD_I_HASH];
        }
#pragma unroll
        for (int b = 0; b < 3; ++b) { st.start[b] = p.start[(size_t)idx * 3 + b]; st.goal[b] = p.goal[(size_t)idx * 3 + b]; }
        const float *ap = p.alpha + (size_t)idx * p.T * 3;
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const int t = lane + 32 * r;
            const bool ok = t < p.T;
            a[r][0] = ok ? ap[t * 3] : 0.0f; a[r][1] = ok ? ap[t * 3 + 1] : 0.0f; a[r][2] = ok ? ap[t * 3 + 2] : 0.0f;
        }
        begin_outer_eval<RPL>(p, lane, st, kind, a, XA);
        return;
    }
}

// End of an inner loop: constraint verdict, lambda escalation, next outer
// iteration or retirement.   optimizer_BLS.py:196-205, optimizer_GD.py:214-224
template <int RPL>
__device__ __forceinline__ void end_inner(const DevParams &p, int lane, Slot &st, int &kind, float (&a)[RPL][3], float4 *XA)
{
    const bool dual = (p.mode == 0) || (p.max_outer > 1);
    bool retire = !dual || st.ful;
    if (!retire) {
        st.lam_sg = st.lam_sg * p.lam_inc; st.lam_jl = st.lam_jl * p.lam_inc;
        st.outer += 1; hash_step(st, 4u);
        retire = st.outer >= p.max_outer;
    }
    if (retire) {
        save_slot<RPL>(p, lane, st, FGD_ST_DONE, a);
        fetch_slot<RPL>(p, lane, st, kind, a, XA);
        return;
    }
    st.inner = 0;
    if (p.mode == 0) st.lr = p.bls_lr0;       // optimizer_BLS.py:193
    begin_outer_eval<RPL>(p, lane, st, kind, a, XA);
}

// Head of the inner loop (optimizer_BLS.py:155-157): continue with a gradient,
// stop at the launch budget, or fall through to the constraint check.
template <int RPL>
__device__ __forceinline__ void inner_head(const DevParams &p, int lane, Slot &st, int &kind, float (&a)[RPL][3], float4 *XA)
{
    if (st.inner < p.max_inner) {
        if (p.budget >= 0 && st.done_iters == p.budget) {
            save_slot<RPL>(p, lane, st, FGD_ST_ACTIVE, a);
            fetch_slot<RPL>(p, lane, st, kind, a, XA);
            return;
        }
        st.done_iters += 1; st.inner_total += 1;
        kind = K_BACK;
        return;
    }
    end_inner<RPL>(p, lane, st, kind, a, XA);
}

// ---------------------------------------------------------------------------
// Persistent optimiser: every warp owns S slots and keeps pulling trajectories
// until the batch queue is empty.  One loop trip = one contraction for every
// slot + the slot's post-processing (candidate evaluation or gradient).
// ---------------------------------------------------------------------------
template <int RPL, int S, bool STRICT, bool KS, int NW>
__global__ void __launch_bounds__(NW * 32) fgd_optimize_kernel(const __grid_constant__ DevParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = p.T, TP = p.TP;
    const SmemLayout L = make_layout(T, TP, p.n_obs, KS, NW * S);
    float *sK = reinterpret_cast<float *>(smem_raw);
    float *sdK = sK + L.k_floats;
    float2 *sObs = reinterpret_cast<float2 *>(sdK + L.k_floats);
    float4 *sX = reinterpret_cast<float4 *>(sObs + L.obs_pairs);
    Slot *sSlot = reinterpret_cast<Slot *>(sX + (size_t)L.n_slots * 2 * L.x_rows);

    if constexpr (KS) {
        const float4 *gK = reinterpret_cast<const float4 *>(p.Kt), *gD = reinterpret_cast<const float4 *>(p.dKt);
        float4 *dK4 = reinterpret_cast<float4 *>(sK), *dD4 = reinterpret_cast<float4 *>(sdK);
        for (int i = threadIdx.x; i < L.k_floats / 4; i += NW * 32) { dK4[i] = __ldg(gK + i); dD4[i] = __ldg(gD + i); }
    }
    for (int i = threadIdx.x; i < p.n_obs; i += NW * 32) sObs[i] = make_float2(p.obs[2 * i], p.obs[2 * i + 1]);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *Kt = KS ? sK : p.Kt, *dKt = KS ? sdK : p.dKt;

    float4 *XA[S], *XB[S];
    Slot *slot[S];
    int kind[S];
    float a[S][RPL][3];     // alpha rows of this lane
    float d[S][RPL][3];     // step direction rows (normalised gradient for BLS, gradient for GD)
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int gs = warp * S + s;
        XA[s] = sX + (size_t)(gs * 2) * L.x_rows;
        XB[s] = XA[s] + L.x_rows;
        slot[s] = sSlot + gs;
        kind[s] = K_IDLE;
#pragma unroll
        for (int r = 0; r < RPL; ++r)
#pragma unroll
            for (int b = 0; b < 3; ++b) { a[s][r][b] = 0.0f; d[s][r][b] = 0.0f; }
        Slot st;
        fetch_slot<RPL>(p, lane, st, kind[s], a[s], XA[s]);
        *slot[s] = st;
    }

    for (;;) {
        bool any = false;
#pragma unroll
        for (int s = 0; s < S; ++s) any |= (kind[s] != K_IDLE);
        if (!any) break;

        const float4 *x1[S], *x2[S];
#pragma unroll
        for (int s = 0; s < S; ++s) { x1[s] = XA[s]; x2[s] = (kind[s] == K_BACK) ? XB[s] : XA[s]; }
        float y1[S][RPL][3], y2[S][RPL][3];
        __syncwarp();
        contract<RPL, S, KS>(Kt, dKt, T, TP, lane, x1, x2, y1, y2);
        __syncwarp();

#pragma unroll
        for (int s = 0; s < S; ++s) {
            if (kind[s] == K_IDLE) continue;
            // warp-uniform scalars: every lane works on its own register copy and
            // all lanes write back identical values (race-free by construction)
            Slot st = *slot[s];
            __syncwarp();
            if (kind[s] == K_BACK) {
                // ---- alpha-gradient, normalisation, first candidate ----------
                float g[RPL][3];
                backward_rows<RPL>(p, y1[s], y2[s], g);
                if (p.mode == 0) {
                    float part = 0.0f;
#pragma unroll
                    for (int r = 0; r < RPL; ++r)
                        if (lane + 32 * r < T) part = part + ss3(g[r][0], g[r][1], g[r][2]);
                    const float nrm = sqrtf(wsum(part));                  // optimizer_BLS.py:165
                    float pb = 0.0f;
#pragma unroll
                    for (int r = 0; r < RPL; ++r) {
#pragma unroll
                        for (int b = 0; b < 3; ++b) d[s][r][b] = g[r][b] / nrm;
                        if (lane + 32 * r < T)
                            pb = pb + ((g[r][0] + g[r][1]) + g[r][2]) * ((d[s][r][0] + d[s][r][1]) + d[s][r][2]);
                    }
                    st.alpha_norm = wsum(pb);                             // optimizer_BLS.py:166
                    st.j = 0;
                } else {
#pragma unroll
                    for (int r = 0; r < RPL; ++r)
#pragma unroll
                        for (int b = 0; b < 3; ++b) d[s][r][b] = g[r][b];
                }
                write_candidate<RPL>(p, lane, st.lr, a[s], d[s], XA[s]);
                kind[s] = K_CAND;
            } else {
                // ---- loss (and, if needed, gradient operands) at alpha or at a candidate
                Rows<RPL> R;
                float loss_c, toc_c;
                int ful_c;
                cost_phase<RPL, STRICT>(p, sObs, lane, y1[s], y2[s], st.start, st.goal, st.lam_sg, st.lam_jl, R, loss_c, toc_c, ful_c);
                if (kind[s] == K_EVAL0) {
                    st.loss = loss_c; st.toc = toc_c; st.ful = ful_c;
                    grad_phase<RPL>(p, lane, R, st.lam_sg, st.lam_jl, XA[s], XB[s]);
                    inner_head<RPL>(p, lane, st, kind[s], a[s], XA[s]);
                } else if (p.mode == 0) {
                    // Armijo test   optimizer_BLS.py:141-149
                    st.cand_evals += 1;
                    const float lr = st.lr, loss = st.loss;
                    const float req = loss - (p.bls_alpha * lr) * st.alpha_norm;
                    if (loss_c > req) {
                        st.lr = lr * p.bls_bm; hash_step(st, 1u);
                        st.j += 1;
                        if (st.j < p.max_bls) {
                            write_candidate<RPL>(p, lane, st.lr, a[s], d[s], XA[s]);
                        } else {
                            // every candidate rejected: new_loss := loss (optimizer_BLS.py:170,178)
                            st.last_new = loss;
                            const bool minimized = (loss - loss < p.eps_loop);
                            if (minimized) { hash_step(st, 3u); end_inner<RPL>(p, lane, st, kind[s], a[s], XA[s]); }
                            else { st.inner += 1; begin_outer_eval<RPL>(p, lane, st, kind[s], a[s], XA[s]); }
                        }
                    } else {
                        accept_candidate<RPL>(p, lr, a[s], d[s]);
                        grad_phase<RPL>(p, lane, R, st.lam_sg, st.lam_jl, XA[s], XB[s]);
                        st.lr = lr * p.bls_bp; st.accepts += 1; hash_step(st, 2u);
                        st.ful = ful_c; st.toc = toc_c; st.last_new = loss_c;
                        const bool minimized = (loss - loss_c < p.eps_loop);     // optimizer_BLS.py:178
                        st.loss = loss_c;
                        if (minimized) { hash_step(st, 3u); end_inner<RPL>(p, lane, st, kind[s], a[s], XA[s]); }
                        else { st.inner += 1; inner_head<RPL>(p, lane, st, kind[s], a[s], XA[s]); }
                    }
                } else {
                    // fixed-step GD   optimizer_GD.py:186-194
                    st.cand_evals += 1;
                    st.last_new = loss_c;
                    if (st.loss - loss_c < p.eps_loop) {
                        hash_step(st, 3u);
                        end_inner<RPL>(p, lane, st, kind[s], a[s], XA[s]);
                    } else {
                        accept_candidate<RPL>(p, st.lr, a[s], d[s]);
                        grad_phase<RPL>(p, lane, R, st.lam_sg, st.lam_jl, XA[s], XB[s]);
                        st.loss = loss_c; st.ful = ful_c; st.toc = toc_c;
                        st.accepts += 1; st.inner += 1; hash_step(st, 2u);
                        inner_head<RPL>(p, lane, st, kind[s], a[s], XA[s]);
                    }
                }
            }
            *slot[s] = st;
        }
    }
}

// ---------------------------------------------------------------------------
// Evaluation only (unit-parity hook and the host's compute_trajectory_cost*):
// one warp per trajectory, grid-stride.
// ---------------------------------------------------------------------------
template <int RPL, bool STRICT, bool KS, int NW>
__global__ void __launch_bounds__(NW * 32) fgd_eval_kernel(const __grid_constant__ DevParams p, const EvalPtrs e)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = p.T, TP = p.TP;
    const SmemLayout L = make_layout(T, TP, p.n_obs, KS, NW);
    float *sK = reinterpret_cast<float *>(smem_raw);
    float *sdK = sK + L.k_floats;
    float2 *sObs = reinterpret_cast<float2 *>(sdK + L.k_floats);
    float4 *sX = reinterpret_cast<float4 *>(sObs + L.obs_pairs);
    if constexpr (KS) {
        const float4 *gK = reinterpret_cast<const float4 *>(p.Kt), *gD = reinterpret_cast<const float4 *>(p.dKt);
        float4 *dK4 = reinterpret_cast<float4 *>(sK), *dD4 = reinterpret_cast<float4 *>(sdK);
        for (int i = threadIdx.x; i < L.k_floats / 4; i += NW * 32) { dK4[i] = __ldg(gK + i); dD4[i] = __ldg(gD + i); }
    }
    for (int i = threadIdx.x; i < p.n_obs; i += NW * 32) sObs[i] = make_float2(p.obs[2 * i], p.obs[2 * i + 1]);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *Kt = KS ? sK : p.Kt, *dKt = KS ? sdK : p.dKt;
    float4 *XA = sX + (size_t)(warp * 2) * L.x_rows, *XB = XA + L.x_rows;

    for (int b = blockIdx.x * NW + warp; b < p.B; b += gridDim.x * NW) {
        const float *ap = p.alpha + (size_t)b * T * 3;
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const int t = lane + 32 * r;
            if (t < T) XA[t] = make_float4(ap[t * 3], ap[t * 3 + 1], ap[t * 3 + 2], 0.0f);
        }
        float start[3], goal[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { start[k] = p.start[(size_t)b * 3 + k]; goal[k] = p.goal[(size_t)b * 3 + k]; }
        const float4 *x1[1] = {XA}, *x2[1] = {XA};
        float y1[1][RPL][3], y2[1][RPL][3];
        __syncwarp();
        contract<RPL, 1, KS>(Kt, dKt, T, TP, lane, x1, x2, y1, y2);
        __syncwarp();
        Rows<RPL> R;
        float loss, toc;
        int ful;
        cost_phase<RPL, STRICT>(p, sObs, lane, y1[0], y2[0], start, goal, e.lam_sg, e.lam_jl, R, loss, toc, ful);
        if (lane == 0) {
            if (e.loss) e.loss[b] = loss;
            if (e.toc) e.toc[b] = toc;
            if (e.fulfilled) e.fulfilled[b] = ful;
        }
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const int t = lane + 32 * r;
            if (t < T) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (e.q) e.q[((size_t)b * T + t) * 3 + k] = R.q[r][k];
                    if (e.v) e.v[((size_t)b * T + t) * 3 + k] = R.v[r][k];
                }
            }
        }
        if (e.grad) {
            grad_phase<RPL>(p, lane, R, e.lam_sg, e.lam_jl, XA, XB);
            const float4 *g1[1] = {XA}, *g2[1] = {XB};
            __syncwarp();
            contract<RPL, 1, KS>(Kt, dKt, T, TP, lane, g1, g2, y1, y2);
            float g[RPL][3];
            backward_rows<RPL>(p, y1[0], y2[0], g);
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int t = lane + 32 * r;
                if (t < T) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) e.grad[((size_t)b * T + t) * 3 + k] = g[r][k];
                }
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// Best restart per problem (random-restart sweep): one warp per problem.
// Key order: constraint-fulfilled first, then lower obstacle cost, then lower index.
// ---------------------------------------------------------------------------
__global__ void fgd_argmin_kernel(int n_problems, int n_restarts, const float *__restrict__ fstate,
                                  const int *__restrict__ istate, int index_offset, float *best_cost, int *best_index)
{
    const int lane = threadIdx.x & 31;
    const int prob = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (prob >= n_problems) return;
    unsigned long long best = ~0ull;
    for (int r = lane; r < n_restarts; r += 32) {
        const size_t b = (size_t)prob * n_restarts + r;
        const float c = fstate[b * FGD_FSTATE + FGD_F_TOC];
        const unsigned ful = istate[b * FGD_ISTATE + FGD_I_FULFILLED] ? 0u : 1u;
        unsigned cb = __float_as_uint(c);
        if (!(c >= 0.0f)) cb = 0x7fffffffu;                 // NaN / negative never wins
        const unsigned long long key = ((unsigned long long)ful << 63) | ((unsigned long long)cb << 31) | (unsigned)r;
        best = key < best ? key : best;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(FULL, best, o);
        best = other < best ? other : best;
    }
    if (lane == 0) {
        const int r = (int)(best & 0x7fffffffu);
        const size_t b = (size_t)prob * n_restarts + r;
        best_cost[prob] = fstate[b * FGD_FSTATE + FGD_F_TOC];
        best_index[prob] = index_offset + (int)b;
    }
}

// ---------------------------------------------------------------------------
// FP32 roofline probe: 8 independent FFMA chains per thread, nothing else.
// MEASURED_PEAKS.json has no FP32 CUDA-core figure, so the harness measures the
// denominator of roofline.frac on the same GPU, in the same run.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fgd_ffma_peak_kernel(int iters, float seed, float *sink)
{
    float a0 = seed + threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f, c = 1e-3f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    const float r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123456.789f) sink[0] = r;
}

}  // namespace fgd
