// fgd_kernels.cuh -- the persistent optimiser kernel, the evaluation kernel and
// the restart-argmin kernel.  See fgd_device.cuh for the mapping.
#pragma once
#include "fgd_device.cuh"
#include "../../include/fgd_b200.h"

#ifdef FGD_DEBUG_TRACE
#include <cstdio>
#define TRACE(...) do { if (blockIdx.x == 0 && threadIdx.x < 32 && (threadIdx.x & 15) == 0) printf(__VA_ARGS__); } while (0)
#else
#define TRACE(...) do { } while (0)
#endif
#ifdef FGD_DEBUG_MARK
#define MARK(id) do { if (p.dbg && blockIdx.x == 0) { volatile int *d_ = p.dbg; d_[threadIdx.x] = (id); } } while (0)
#else
#define MARK(id) do { } while (0)
#endif

#ifdef FGD_PHASE_CLOCKS      // debug builds: where does a lone team spend its cycles? (CTA 0, team 0 -> p.dbg)
#define PCLK(i) do { const long long c_ = clock64(); pc[i] += c_ - pc_t; pc_t = c_; } while (0)
#else
#define PCLK(i) do { } while (0)
#endif

namespace fgd {

// shared-memory carve-up, identical on host and device
struct SmemLayout {
    int k_floats;     // T*TP when the K tables are staged in shared memory (KD = 2x, KO = 1x), 0 when they stay in L2
    int obs_pairs;    // obstacle slots of one copy of the obstacle set (padded to even)
    int obs_copies;   // 1: one copy per CTA; n_teams: a private copy per team (live obstacle updates)
    int x_rows;       // float4 rows per operand buffer
    int n_teams;      // trajectory teams per CTA
    int xch_words;    // exchange scratch words per team (0 for single-warp teams)
    __host__ __device__ size_t team_bytes() const { return (size_t)2 * x_rows * 16 + (size_t)xch_words * 4; }
    __host__ __device__ size_t bytes() const { return (size_t)3 * k_floats * 4 + (size_t)obs_copies * obs_pairs * 8 + (size_t)n_teams * team_bytes(); }
};

// n_obs: obstacles of the launch; live launches size every team's private copy for the capacity instead
__host__ __device__ inline SmemLayout make_layout(int T, int TP, int n_obs, int ksrc, int n_teams, int wpt, bool per_team_obs = false)
{
    SmemLayout l;
    l.k_floats = ksrc == K_SMEM ? T * TP : 0;
    l.obs_pairs = (n_obs + 2) & ~1;
    l.obs_copies = per_team_obs ? n_teams : 1;
    l.x_rows = T | 1;     // odd row count: the operand buffers of neighbouring teams start 4 banks apart (mod 8)
    l.n_teams = n_teams;
    l.xch_words = wpt > 1 ? XCH_WORDS : 0;
    return l;
}

__device__ __forceinline__ void hash_step(Slot &st, unsigned code) { st.hash = st.hash * 1000003u + code; }

// rows of this thread (alpha, a candidate, ...) into an operand buffer
template <int WPT>
__device__ __forceinline__ void write_rows(const int T, const Team<WPT> &G, const f2 (&c)[3], float4 *X)
{
    const int t = G.tl * R;
    if (t < T) X[t] = make_float4(c[0].x, c[1].x, c[2].x, 0.0f);
    if (t + 1 < T) X[t + 1] = make_float4(c[0].y, c[1].y, c[2].y, 0.0f);
}

// candidate  (1 - lam_reg*lr) * alpha - lr * dir      optimizer_BLS.py:139, optimizer_GD.py:185
__device__ __forceinline__ void make_candidate(const DevParams &p, float lr, const f2 (&a)[3], const f2 (&d)[3], f2 (&c)[3])
{
    const float c1 = 1.0f - p.lam_reg * lr;
#pragma unroll
    for (int b = 0; b < 3; ++b) c[b] = fma2(bc2(c1), a[b], neg2(mul2(bc2(lr), d[b])));
}

// Trajectory I/O goes through the team's operand buffer XA viewed as 3T packed floats ("stage"), so that global
// memory only sees linear, fully coalesced 128 B accesses.  That matters most for the zero-copy host path
// (fgd_optimize_host_io on pinned buffers): p.alpha_in / p.alpha / p.start / p.goal / p.fstate / p.istate are then
// mapped HOST pointers and every access is a PCIe transaction - the kernel pulls a trajectory when a team picks it
// up and pushes the result when the team retires it, overlapped with the other teams' arithmetic.
template <int WPT>
__device__ __forceinline__ void save_slot(const DevParams &p, const Team<WPT> &G, const Slot &st, int status, const f2 (&a)[3], float *stage)
{
    const int b = st.traj;
    const int t = G.tl * R;
    G.sync();                                         // the team is done with whatever XA held
    if (t < p.T) { stage[t * 3] = a[0].x; stage[t * 3 + 1] = a[1].x; stage[t * 3 + 2] = a[2].x; }
    if (t + 1 < p.T) { stage[t * 3 + 3] = a[0].y; stage[t * 3 + 4] = a[1].y; stage[t * 3 + 5] = a[2].y; }
    G.sync();
    float *ap = p.alpha + (size_t)b * p.T * 3;
    const int n = p.T * 3;
    for (int i = G.tl; i < n; i += WPT * 32) ap[i] = stage[i];
    if (G.tl < FGD_FSTATE) {                          // one 32 B row of float state, one of integer state
        const int i = G.tl;
        const float fv = i == FGD_F_LAM_SG ? st.lam_sg : i == FGD_F_LAM_JL ? st.lam_jl : i == FGD_F_LR ? st.lr
                       : i == FGD_F_LOSS ? st.loss : i == FGD_F_TOC ? st.toc : i == FGD_F_LAST_NEW_LOSS ? st.last_new : 0.0f;
        const int iv = i == FGD_I_STATUS ? status : i == FGD_I_OUTER ? st.outer : i == FGD_I_INNER ? st.inner
                     : i == FGD_I_INNER_TOTAL ? st.inner_total : i == FGD_I_CAND_EVALS ? st.cand_evals
                     : i == FGD_I_ACCEPTS ? st.accepts : i == FGD_I_FULFILLED ? st.ful : (int)st.hash;
        p.fstate[(size_t)b * FGD_FSTATE + i] = fv;
        p.istate[(size_t)b * FGD_ISTATE + i] = iv;
    }
    G.sync();                                         // stage is free again
}

// Speculative line search (SPEC replicas of one trajectory per CTA): what the replicas exchange per candidate round.
struct SpecScratch {
    float loss[8], toc[8];
    int ful[8], flag[8];       // flag: 0 = this candidate passes the Armijo test, 1 = rejected, 2 = beyond max_bls_iteration
    unsigned nz[R];            // non-zero rows of the winner's velocity-gradient operand
    unsigned fetch;            // queue index of the trajectory all replicas pick up
};

// Pull the next unfinished trajectory (team-uniform).  The first pick of every team is static - team t of CTA c takes
// trajectory t * gridDim + c - so that a batch smaller than the resident teams spreads over all SMs instead of filling the
// first CTAs; later picks come from the global queue counter, which starts behind the static round.
// SPEC: all warps of the CTA are replicas of ONE team; warp 0 picks, everybody loads.
template <int WPT, int MODE, int TEAMS, bool SP>
__device__ __forceinline__ void fetch_slot(const DevParams &p, const Team<WPT> &G, Slot &st, int &kind, f2 (&a)[3], float *stage, bool &first,
                                           SpecScratch *sp)
{
    constexpr int TEAMS_Q = SP ? 1 : TEAMS;
    for (;;) {
        unsigned idx = 0;
        if (first) {
            first = false;
            idx = (SP ? 0u : (unsigned)((threadIdx.x >> 5) / WPT)) * gridDim.x + blockIdx.x;
        } else if constexpr (SP) {
            __syncthreads();                          // the previous round's readers are done
            if (threadIdx.x == 0) sp->fetch = (unsigned)TEAMS_Q * gridDim.x + atomicAdd(p.queue, 1u);
            __syncthreads();
            idx = sp->fetch;
        } else if constexpr (WPT == 1) {
            if (G.lane == 0) idx = (unsigned)TEAMS_Q * gridDim.x + atomicAdd(p.queue, 1u);
            idx = __shfl_sync(FULL, idx, 0);
        } else {
            unsigned *xf = reinterpret_cast<unsigned *>(G.xch + XCH_FETCH);
            G.sync();                                 // the previous round's readers are done
            if (G.tl == 0) *xf = (unsigned)TEAMS_Q * gridDim.x + atomicAdd(p.queue, 1u);
            G.sync();
            idx = *xf;
        }
        if (idx >= (unsigned)p.B) { st.traj = -1; kind = K_IDLE; return; }
        const int *is = p.istate + (size_t)idx * FGD_ISTATE;
        const int status = p.fresh ? FGD_ST_FRESH : is[FGD_I_STATUS];
        if (status == FGD_ST_DONE) continue;          // finished in an earlier launch, take the next one
        const float *fs = p.fstate + (size_t)idx * FGD_FSTATE;
        st.traj = (int)idx;
        st.done_iters = 0; st.j = 0; st.alpha_norm = 0.0f;
        if (status == FGD_ST_FRESH) {
            st.lam_sg = p.lam_sg0; st.lam_jl = p.lam_jl0;
            st.lr = (MODE == 0) ? p.bls_lr0 : p.gd_lr[0];
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
        const float *ap = p.alpha_in + (size_t)idx * p.T * 3;
        const int n = p.T * 3;
        for (int i = G.tl; i < n; i += WPT * 32) stage[i] = ap[i];
        G.sync();
        const int t = G.tl * R;
        const bool ok = t < p.T, ok1 = t + 1 < p.T;
#pragma unroll
        for (int c = 0; c < 3; ++c) a[c] = mk2(ok ? stage[t * 3 + c] : 0.0f, ok1 ? stage[t * 3 + 3 + c] : 0.0f);
        G.sync();                                     // rows are in registers: the caller may overwrite XA
        kind = K_EVAL0;
        return;
    }
}

// K tables -> shared memory with the TMA bulk-copy engine (cp.async.bulk, SASS UBLKCP): one thread
// programs two copies (KD, KO) that complete on an mbarrier; every thread of the CTA then waits on the
// barrier's phase 0.  The obstacle set (8 B granules, any count) is staged by the threads meanwhile.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int KS>
__device__ __forceinline__ void stage_constants(const DevParams &p, const SmemLayout &L, float *sKD, float *sKO, float2 *sObs, int nthreads,
                                                bool stage_obs = true)
{
    __shared__ __align__(8) unsigned long long tma_bar;
    if constexpr (KS == K_SMEM) {
        const unsigned bar = smem_u32(&tma_bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned bytes_kd = (unsigned)(2 * L.k_floats) * 4u, bytes_ko = (unsigned)L.k_floats * 4u;     // multiples of 256 B
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes_kd + bytes_ko) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sKD)), "l"(p.KD), "r"(bytes_kd), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sKO)), "l"(p.KO), "r"(bytes_ko), "r"(bar) : "memory");
        }
    }
    // trip counts are CTA-uniform (the bound check is inside), so no warp diverges ahead of the barrier
#pragma unroll 1
    for (int i0 = 0; stage_obs && i0 < L.obs_pairs; i0 += nthreads) {
        const int i = i0 + threadIdx.x;
        if (i < L.obs_pairs) sObs[i] = (i < p.n_obs) ? make_float2(p.obs[2 * i], p.obs[2 * i + 1]) : make_float2(0.f, 0.f);
    }
    if constexpr (KS == K_SMEM) {
        const unsigned bar = smem_u32(&tma_bar);
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(0u) : "memory");
        }
    }
    __syncthreads();
}

// K / dK operand table -> tensor memory (see fgd_device.cuh): warp 0 allocates COLS columns, warps 0..3 write the table
// into their lane quadrant, every warp gets the address of the quadrant it may read.  Quadrant q holds the entries of team
// thread 32 (q % WPT) + lane: single-warp teams (WPT = 1) - the same 32 threads in every quadrant; multi-warp teams - warp
// w of the CTA is warp w % WPT of its team.  CPK = 4: lane columns 4k..4k+3 = KD[k][thread] (K and dK of the thread's two
// rows); CPK = 2: columns 2k, 2k+1 = KO[k][thread] (K alone).  CTA-uniform control flow; call once, before the main loop.
template <int WPT, int CPK, int COLS>
__device__ __forceinline__ unsigned tmem_stage_tables(const DevParams &p)
{
    __shared__ unsigned tm_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tm_base)), "r"((unsigned)COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tk = tm_base + ((unsigned)((warp & 3) * 32) << 16);
    if (warp < 4) {
        constexpr int NT = WPT * 32;                                            // threads per team = table entries per k
        const int thr = (warp & (WPT - 1)) * 32 + lane;
        // batches of 8 columns: the loads of a batch are in flight together (a one-column loop exposes the L2 / HBM latency
        // T times: ~15 us of every launch's prologue at T = 50)
        constexpr int KB = 8;
#pragma unroll 1
        for (int k0 = 0; k0 < p.T; k0 += KB) {
            if constexpr (CPK == 4) {
                const float4 *src = reinterpret_cast<const float4 *>(p.KD) + thr;   // KD[k][thread][4]
                float4 v[KB];
#pragma unroll
                for (int u = 0; u < KB; ++u) v[u] = (k0 + u < p.T) ? __ldg(src + (size_t)(k0 + u) * NT) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < KB; ++u)
                    if (k0 + u < p.T)
                        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tk + 4 * (k0 + u)), "r"(__float_as_uint(v[u].x)),
                                     "r"(__float_as_uint(v[u].y)), "r"(__float_as_uint(v[u].z)), "r"(__float_as_uint(v[u].w)) : "memory");
            } else {
                const float2 *src = reinterpret_cast<const float2 *>(p.KO) + thr;   // KO[k][thread][2]
                float2 v[KB];
#pragma unroll
                for (int u = 0; u < KB; ++u) v[u] = (k0 + u < p.T) ? __ldg(src + (size_t)(k0 + u) * NT) : make_float2(0.f, 0.f);
#pragma unroll
                for (int u = 0; u < KB; ++u)
                    if (k0 + u < p.T)
                        asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(tk + 2 * (k0 + u)), "r"(__float_as_uint(v[u].x)),
                                     "r"(__float_as_uint(v[u].y)) : "memory");
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return tk;
}

// every warp of the CTA is done with the tables: give the columns back
template <int COLS>
__device__ __forceinline__ void tmem_release_tables(unsigned tk)
{
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tk & 0x0000ffffu), "r"((unsigned)COLS) : "memory");
}

// ---------------------------------------------------------------------------
// Live obstacle updates (fgd_optimize_live; single-warp teams).  The host publishes obstacle generation g into ring slot
// g % FGD_OBS_RING with stream-ordered copies: slot header := -1, data, slot header := (g, count), latest := g.  A team
// adopts the latest generation seqlock-style: header == g before and after its copy into the team's private shared-memory
// set, otherwise it retries with whatever is latest by then.  All ring reads bypass L1 (a slot is rewritten while the
// kernel runs).  Returns true when the team's set changed.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire_sys(const int *p)
{
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float2 ld_volatile_f2(const float2 *p)
{
    float2 v;
    asm volatile("ld.volatile.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ bool live_refresh(const DevParams &p, const int lane, float2 *sObsT, int &cur_gen, int &n_obs)
{
    bool changed = false;
    for (int tries = 0; tries < 65536; ++tries) {             // bounded: a publisher that died mid-update must not hang the kernel
        int gen = 0, hdr = 0, cnt = 0;
        if (lane == 0) gen = ld_acquire_sys(p.obs_meta);
        gen = __shfl_sync(FULL, gen, 0);
        if (gen == cur_gen) return changed;
        const int s = gen & (FGD_OBS_RING - 1);
        if (lane == 0) { hdr = ld_acquire_sys(p.obs_meta + 2 + 2 * s); cnt = ld_acquire_sys(p.obs_meta + 3 + 2 * s); }
        hdr = __shfl_sync(FULL, hdr, 0); cnt = __shfl_sync(FULL, cnt, 0);
        if (hdr != gen || cnt < 0 || cnt > p.obs_cap) continue;        // the slot is being rewritten: a newer generation is on its way
        const float2 *src = reinterpret_cast<const float2 *>(p.obs_ring) + (size_t)s * p.obs_cap;
        const int pairs = (cnt + 2) & ~1;
        __syncwarp();                                                 // nobody of the team still reads the old set
        for (int i = lane; i < pairs; i += 32) sObsT[i] = i < cnt ? ld_volatile_f2(src + i) : make_float2(0.f, 0.f);
        __syncwarp();                                                 // every lane's ring reads have returned (their values are in shared memory)
        if (lane == 0) hdr = ld_acquire_sys(p.obs_meta + 2 + 2 * s);
        hdr = __shfl_sync(FULL, hdr, 0);
        if (hdr != gen) continue;                                      // overwritten while copying
        cur_gen = gen; n_obs = cnt; changed = true;
        return true;
    }
    if (cur_gen < 0) n_obs = 0;                                // never got a consistent set: run without obstacles rather than on garbage
    return changed;
}

// switch log of trajectory b: entry 0 = (count, 0), entries 1..count = (inner iterations completed, generation adopted)
__device__ __forceinline__ void log_switch(const DevParams &p, const int lane, const int traj, int &n_switch, const int inner_total, const int gen)
{
    n_switch += 1;
    if (p.switch_log && lane == 0) {
        int *row = p.switch_log + (size_t)traj * FGD_SWITCH_LOG * 2;
        row[0] = n_switch; row[1] = 0;
        if (n_switch < FGD_SWITCH_LOG) { row[2 * n_switch] = inner_total; row[2 * n_switch + 1] = gen; }
    }
}

// ---------------------------------------------------------------------------
// Persistent optimiser: every team of WPT warps runs one trajectory as an
// autonomous state machine and keeps pulling trajectories until the batch queue
// is empty.  One loop trip = one contraction + the post-processing of its result
// (candidate evaluation or gradient).  NW warps per CTA; WPT > 1 requires
// NW == WPT (the CTA barrier is the team barrier).
// ---------------------------------------------------------------------------
// TC > 0: instance specialised for T == TC (the reference's default T = 50): the contraction loops are fully unrolled.
// MODE: 0 = backtracking line search (optimizer_BLS.py), 1 = gradient descent (optimizer_GD.py) - a compile-time
// constant so that each instance carries only its own state machine.
// SPEC > 0 (BLS, single-warp teams, NW == SPEC): latency mode for batches smaller than the machine.  The CTA runs ONE
// trajectory; its SPEC warps are replicas that hold the same state and evaluate the Armijo candidates j, j+1, ..., j+SPEC-1
// of a line search AT THE SAME TIME (north_star item 5; the reference tries them one after the other,
// optimizer_BLS.py:131-150).  The first accepting candidate in the reference's order wins; counters, hash and step size
// are advanced exactly as the sequential search would have - candidates evaluated beyond the winner leave no trace - so
// the iterates are the reference's bit for bit.  Gradient trips are computed redundantly by all replicas from the
// winner's operand buffer.
// LIVE: live obstacle updates (fgd_optimize_live): a private obstacle set per team, polled and refreshed in the kernel.
// HELP: the obstacle loop of a many-obstacle scene is shared with the warp's sample-less lanes (fgd_device.cuh, share_split);
// the host launches this instance when the split is active, the LIVE instances always carry it.
// OC > 0: instance for exactly OC obstacles (the default scene's 11, next to TC = 50): unrolled obstacle loop.
template <int WPT, bool STRICT, int KS, int NW, int MINB, bool ARM, int TC, int MODE, int SPEC = 0, bool LIVE = false, bool HELP = LIVE, int OC = 0>
__global__ void __launch_bounds__(NW * 32, MINB) fgd_optimize_kernel(const __grid_constant__ DevParams p)
{
    constexpr bool SP = SPEC > 0;
    static_assert(OC == 0 || (TC > 0 && !LIVE && !HELP && !ARM && WPT == 1 && OC < FGD_SHARE_MIN_OBS), "compile-time obstacle count: the T = TC default instance");
    static_assert(!HELP || (WPT == 1 && !ARM), "helper lanes: single-warp teams, end-effector cost");
    static_assert(!LIVE || (WPT == 1 && !SP), "live obstacle updates: single-warp teams");
    static_assert(!SP || (WPT == 1 && MODE == 0 && NW == SPEC && SPEC <= 8), "speculative line search: BLS, single-warp replicas, one trajectory per CTA");
    static_assert(TC == 0 || KS == K_TMEM, "compile-time T: TMEM instances only");
    static_assert(WPT == 1 || NW == WPT || KS == K_TMEM, "multi-warp teams own their CTA unless they share the tensor-memory tables");
    static_assert(KS != K_TMEM || (NW >= 4 && NW % 4 == 0 && NW % WPT == 0 && (WPT == 1 || NW / WPT <= 15)), "TMEM tables: whole lane quadrants, whole teams, one named barrier per team");
    // tensor-memory columns per k and in total: single-warp teams K and dK of T <= 64 (two CTAs per SM may be resident);
    // multi-warp teams own the SM: WPT = 2 K and dK of T <= 128, WPT = 4 K alone of T <= 256
    constexpr int CPK = (KS == K_TMEM && WPT == 4) ? 2 : 4;
    constexpr int TM_COLS = WPT == 1 ? TMEM_COLS : 512;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TEAMS = NW / WPT;
    const int T = TC > 0 ? TC : p.T;
    constexpr bool live = LIVE;
    const SmemLayout L = make_layout(T, WPT * 32 * R, live ? p.obs_cap : p.n_obs, KS, TEAMS, WPT, live);
    float *sKD = reinterpret_cast<float *>(smem_raw);
    float *sKO = sKD + 2 * L.k_floats;
    float2 *sObs = reinterpret_cast<float2 *>(sKO + L.k_floats);
    unsigned char *sTeams = reinterpret_cast<unsigned char *>(sObs + (size_t)L.obs_copies * L.obs_pairs);
    stage_constants<KS>(p, L, sKD, sKO, sObs, NW * 32, !live);

    const int team = (threadIdx.x >> 5) / WPT;
    if (live) sObs += (size_t)team * L.obs_pairs;
    float4 *XA = reinterpret_cast<float4 *>(sTeams + (size_t)team * L.team_bytes()), *XB = XA + L.x_rows;
    const Team<WPT> G(reinterpret_cast<float *>(XB + L.x_rows), (WPT > 1 && NW > WPT) ? 1 + team : 0);
    const float *kd = (KS == K_SMEM ? sKD : p.KD) + G.tl * 2 * R;
    const float *ko = (KS == K_SMEM ? sKO : p.KO) + G.tl * R;
    const float *dd = p.DO + G.tl * R;
    unsigned tk = 0;
    if constexpr (KS == K_TMEM) tk = tmem_stage_tables<WPT, CPK, TM_COLS>(p);
    (void)kd; (void)ko; (void)dd; (void)tk;
    int n_obs_live = 0, cur_gen = -1, polled_at = -1, n_switch = 0;       // LIVE: the team's obstacle set and poll bookkeeping
    (void)n_obs_live; (void)cur_gen; (void)polled_at; (void)n_switch;
    // two-chain obstacle sums (fgd_device.cuh, share_split): 2 = on the helper lanes, 1 = both chains in the owner lane, 0 = never
    // active here (the default single-warp-team instance: the host launches its HELP twin for every scene with an active split)
    constexpr int SHARE = HELP ? 2 : ((WPT == 1 && KS == K_TMEM && !ARM && !SP && !LIVE) ? 0 : 1);
    int split = (LIVE || SHARE == 0) ? 0 : share_split(T, p.n_obs, ARM);  // LIVE: follows the adopted obstacle set
    if constexpr (SHARE == 0) {
        if (share_split(T, p.n_obs, ARM) > 0) __trap();                    // launched for a scene that needs the HELP twin: a host-side dispatch bug
    }
    if constexpr (OC > 0) {
        if (p.n_obs != OC) __trap();                                       // the OC instance is only valid for exactly OC obstacles
    }
    __shared__ SpecScratch sp_mem;                                          // SPEC only (a few words)
    SpecScratch *sp = &sp_mem;
    const int rep = SP ? (int)(threadIdx.x >> 5) : 0;                       // replica index = candidate offset in a round
    int win = rep;                                                          // replica whose operand buffers feed the next gradient trip
    float lr_mine = 0.0f;                                                   // step size of this replica's candidate
    auto gsync = [&]() { if constexpr (SP) __syncthreads(); else G.sync(); };
    bool first_fetch = true;
    (void)win; (void)lr_mine;

    int kind = K_IDLE;
    Slot st;                   // team-uniform loop state of this team's trajectory (registers)
    st.traj = -1;
    unsigned nz[WPT][R];       // non-zero rows of the velocity gradient operand (see contract_back)
#pragma unroll
    for (int w = 0; w < WPT; ++w) { nz[w][0] = 0u; nz[w][1] = 0u; }
    f2 a[3], d[3];             // alpha rows and step direction rows (normalised gradient for BLS, gradient for GD)
#pragma unroll
    for (int b = 0; b < 3; ++b) { a[b] = bc2(0.0f); d[b] = bc2(0.0f); }
    bool boot = true;          // first trip: nothing to contract yet, just fill the slot through the common tail
#ifdef FGD_PHASE_CLOCKS
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pc_t = clock64();
#endif
    for (;;) {
        f2 y1[3], y2[3];
        const bool was_back = (kind == K_BACK);
        (void)was_back;
        if (!boot) {
            if (kind == K_IDLE) break;
            gsync();                                                                     // operands complete
        }
        TRACE("lane %d trip boot=%d kind=%d\n", threadIdx.x, (int)boot, kind);
        bool want_cand = false;     // write the next candidate into XA
        bool want_head = false;     // go to the head of the inner loop
        bool want_end = false;      // inner loop finished: constraint check / lambda escalation
        bool want_eval = false;     // (re-)evaluate loss and gradient operands at alpha

        if (!boot && kind == K_BACK) {
            // ---- backward contraction: K G_q + dK (-G_v) ---------------------------------------
            if constexpr (KS == K_TMEM && WPT > 1) {                                             // multi-warp teams sharing the SM's tensor memory
                load_nz<WPT>(G, nz);
                contract_back_mw<WPT, CPK>(tk, dd, T, XA, XB, nz, y1, y2);
            } else if constexpr (KS == K_TMEM) {
                if constexpr (SP) {                                                              // the winner's gradient operands
                    const float4 *XAw = reinterpret_cast<const float4 *>(sTeams + (size_t)win * L.team_bytes());
                    if (win != rep) { nz[0][0] = sp->nz[0]; nz[0][1] = sp->nz[1]; }
                    contract_back_tm<TC>(tk, T, XAw, XAw + L.x_rows, nz, y1, y2);
                } else contract_back_tm<TC>(tk, T, XA, XB, nz, y1, y2);
            } else {
                load_nz<WPT>(G, nz);
                contract_back<WPT, KS>(ko, kd, T, XA, XB, nz, y1, y2);
            }
            gsync();                                                                     // operands consumed
            PCLK(4);
            // ---- alpha-gradient, normalisation, first candidate ------------------------------
            f2 g[3];
            backward_rows(p, y1, y2, g);
            float alpha_norm = 0.0f;
            if (MODE == 0) {
                const bool v0 = G.tl * R < T, v1 = G.tl * R + 1 < T;
                float part = 0.0f;
                const f2 ss = ss3_2(g[0], g[1], g[2]);
                if (v0) part = part + ss.x;
                if (v1) part = part + ss.y;
                const float scale = 1.0f / sqrtf(tsum<WPT>(G, part, XCH_NORM));          // optimizer_BLS.py:165 (MUFU.RSQ measured: < 1 % gain, not taken)
                float pb = 0.0f;
                const f2 n0 = mul2(g[0], bc2(scale)), n1 = mul2(g[1], bc2(scale)), n2 = mul2(g[2], bc2(scale));
                const f2 pp = mul2(add2(add2(g[0], g[1]), g[2]), add2(add2(n0, n1), n2));
                if (v0) pb = pb + pp.x;
                if (v1) pb = pb + pp.y;
                g[0] = n0; g[1] = n1; g[2] = n2;
                alpha_norm = tsum<WPT>(G, pb, XCH_ANORM);                                // optimizer_BLS.py:166
            }
#pragma unroll
            for (int b = 0; b < 3; ++b) d[b] = g[b];
            st.alpha_norm = alpha_norm; st.j = 0;
            want_cand = true;
            kind = K_CAND;
        } else if (!boot) {
            // ---- forward contraction: K x, dK x for x = alpha or a candidate -------------------
            if constexpr (KS == K_TMEM && WPT > 1) {
                if constexpr (CPK == 4) contract_tm<true, 0>(tk, T, XA, XA, y1, y2);
                else contract_tm_k(tk, dd, WPT * 32 * R, T, XA, y1, y2);
            } else if constexpr (KS == K_TMEM) {
                contract_tm<true, TC>(tk, T, XA, XA, y1, y2);
            } else {
                contract<WPT, KS, true>(kd, T, XA, XA, y1, y2);
            }
            gsync();                                                                     // operands consumed
            PCLK(0);
            // ---- loss (and, if accepted, gradient operands) at alpha or at a candidate -------
            Rows<ARM> Rw;
            float loss_c, toc_c;
            int ful_c;
            cost_phase<WPT, STRICT, ARM, (WPT > 1), LIVE, SHARE, OC>(p, T, sObs, LIVE ? n_obs_live : p.n_obs, G, y1, y2, st.start, st.goal, st.lam_sg, st.lam_jl, Rw, loss_c, toc_c, ful_c,
                                                                split, XA, XB);
            PCLK(1);
            bool accept = false;
            if (kind == K_EVAL0) {
                st.loss = loss_c; st.toc = toc_c; st.ful = ful_c;
                accept = true; want_head = true;
                win = rep;
            } else if constexpr (SP) {
                // ---- one round of the speculative line search: replica r holds candidate j + r ----
                const float loss = st.loss;
                const float req = loss - (p.bls_alpha * lr_mine) * st.alpha_norm;        // Armijo test, optimizer_BLS.py:141-149
                const int flag = (st.j + rep >= p.max_bls) ? 2 : ((loss_c > req) ? 1 : 0);
                if (G.lane == 0) { sp->loss[rep] = loss_c; sp->toc[rep] = toc_c; sp->ful[rep] = ful_c; sp->flag[rep] = flag; }
                __syncthreads();
                int acc = -1, n_rej = 0;
                bool stop = false;
#pragma unroll
                for (int r = 0; r < SPEC; ++r) {
                    const int f = sp->flag[r];
                    if (!stop) { if (f == 1) n_rej += 1; else { stop = true; if (f == 0) acc = r; } }
                }
                for (int i = 0; i < n_rej; ++i) {          // the rejected candidates the sequential search went through
                    st.lr = st.lr * p.bls_bm; hash_step(st, 1u);
                    st.j += 1; st.cand_evals += 1;
                }
                if (acc >= 0) {
                    st.cand_evals += 1;
                    const float lr = st.lr, lc = sp->loss[acc];
                    const bool minimized = (loss - lc < p.eps_loop);                   // optimizer_BLS.py:178
                    f2 c[3];
                    make_candidate(p, lr, a, d, c);             // the winner's candidate, recomputed by every replica: same bits
#pragma unroll
                    for (int b = 0; b < 3; ++b) a[b] = c[b];
                    st.lr = lr * p.bls_bp;
                    st.accepts += 1; hash_step(st, 2u);
                    st.ful = sp->ful[acc]; st.toc = sp->toc[acc]; st.last_new = lc; st.loss = lc;
                    if (minimized) { hash_step(st, 3u); want_end = true; }
                    else { st.inner += 1; want_head = true; }
                    win = acc;
                    accept = (rep == acc);                      // only the winner holds the rows of the accepted candidate
                } else if (st.j < p.max_bls) {
                    want_cand = true;
                } else {
                    st.last_new = loss;                         // every candidate rejected: new_loss := loss (optimizer_BLS.py:170,178)
                    if (loss - loss < p.eps_loop) { hash_step(st, 3u); want_end = true; }
                    else { st.inner += 1; want_eval = true; }
                }
            } else {
                st.cand_evals += 1;
                const float lr = st.lr, loss = st.loss;
                const bool minimized = (loss - loss_c < p.eps_loop);               // optimizer_BLS.py:178, optimizer_GD.py:194
                bool rejected = false;
                if (MODE == 0) {                                                   // Armijo test, optimizer_BLS.py:141-149
                    const float req = loss - (p.bls_alpha * lr) * st.alpha_norm;
                    rejected = loss_c > req;
                } else {
                    st.last_new = loss_c;
                }
                if (rejected) {
                    st.lr = lr * p.bls_bm; hash_step(st, 1u);
                    st.j += 1;
                    if (st.j < p.max_bls) {
                        want_cand = true;
                    } else {
                        // every candidate rejected: new_loss := loss (optimizer_BLS.py:170,178)
                        st.last_new = loss;
                        if (loss - loss < p.eps_loop) { hash_step(st, 3u); want_end = true; }
                        else { st.inner += 1; want_eval = true; }
                    }
                } else if (MODE == 1 && minimized) {
                    hash_step(st, 3u); want_end = true;    // GD: the candidate is discarded (optimizer_GD.py:191-192)
                } else {
                    f2 c[3];
                    make_candidate(p, lr, a, d, c);
#pragma unroll
                    for (int b = 0; b < 3; ++b) a[b] = c[b];
                    accept = true;
                    if (MODE == 0) st.lr = lr * p.bls_bp;
                    st.accepts += 1; hash_step(st, 2u);
                    st.ful = ful_c; st.toc = toc_c; st.last_new = loss_c; st.loss = loss_c;
                    if (minimized) { hash_step(st, 3u); want_end = true; }       // BLS keeps the accepted alpha
                    else { st.inner += 1; want_head = true; }
                }
            }
            if (accept) grad_phase<WPT, ARM>(p, T, G, Rw, st.start, st.goal, st.lam_sg, st.lam_jl, XA, XB, nz);
            if constexpr (SP) {
                if (accept && kind == K_CAND && G.lane == 0) { sp->nz[0] = nz[0][0]; sp->nz[1] = nz[0][1]; }
            }
        }
        PCLK(was_back ? 5 : 2);
        // ---- common tail: loop heads, retirement, refill -------------------------------------
        bool save_active = false;
        if (want_head) {                                         // optimizer_BLS.py:155-157
            if (st.inner < p.max_inner) {
                bool switched = false;
                if constexpr (LIVE) {
                    // live obstacle updates: every poll_every inner iterations of this trajectory look for a newer obstacle set
                    // (plain-loop semantics, optimizer_BLS.py:79,82,90: the environment is re-read between iterations)
                    if (st.inner_total != polled_at && st.inner_total % p.poll_every == 0) {
                        polled_at = st.inner_total;
                        switched = live_refresh(p, G.lane, sObs, cur_gen, n_obs_live);
                        split = share_split(T, n_obs_live, false);
                        if (switched) log_switch(p, G.lane, st.traj, n_switch, st.inner_total, cur_gen);
                    }
                }
                if (switched) want_eval = true;                  // like a resumed launch: the loss at the new obstacle set first
                else if (p.budget >= 0 && st.done_iters == p.budget) save_active = true;
                else { st.done_iters += 1; st.inner_total += 1; kind = K_BACK; }
            } else {
                want_end = true;
            }
        }
        bool retire = false;
        if (want_end) {                                          // optimizer_BLS.py:196-205, optimizer_GD.py:214-224
            const bool dual = (MODE == 0) || (p.max_outer > 1);
            retire = !dual || st.ful;
            if (!retire) {
                st.lam_sg = st.lam_sg * p.lam_inc; st.lam_jl = st.lam_jl * p.lam_inc;
                st.outer += 1; hash_step(st, 4u);
                retire = st.outer >= p.max_outer;
            }
            if (!retire) {
                st.inner = 0;
                if (MODE == 0) st.lr = p.bls_lr0;              // optimizer_BLS.py:193
                want_eval = true;
            }
        }
        if (retire || save_active || boot) {
            if (!boot && rep == 0) save_slot<WPT>(p, G, st, retire ? FGD_ST_DONE : FGD_ST_ACTIVE, a, reinterpret_cast<float *>(XA));
            fetch_slot<WPT, MODE, TEAMS, SP>(p, G, st, kind, a, reinterpret_cast<float *>(XA), first_fetch, sp);
            want_eval = (kind != K_IDLE);
            if constexpr (LIVE) {
                if (kind != K_IDLE) {                          // a trajectory starts with the latest obstacle set
                    live_refresh(p, G.lane, sObs, cur_gen, n_obs_live);
                    split = share_split(T, n_obs_live, false);
                    n_switch = 0; polled_at = st.inner_total;
                    log_switch(p, G.lane, st.traj, n_switch, st.inner_total, cur_gen);
                }
            }
        }
        if (want_eval) {
            // (re)start with the loss and gradient operands at the current alpha: optimizer_BLS.py:163, optimizer_GD.py:210
            if (MODE == 1) st.lr = p.gd_lr[st.outer];
            write_rows<WPT>(T, G, a, XA);
            kind = K_EVAL0;
        }
        if (want_cand) {
            f2 c[3];
            lr_mine = st.lr;
            if constexpr (SP) {
                for (int i = 0; i < rep; ++i) lr_mine = lr_mine * p.bls_bm;      // the step the sequential search would try rep rejections later
            }
            make_candidate(p, lr_mine, a, d, c);
            write_rows<WPT>(T, G, c, XA);
        }
        boot = false;
        PCLK(was_back ? 6 : 3);
#ifdef FGD_PHASE_CLOCKS
        pc[7] += 1;
#endif
    }
#ifdef FGD_PHASE_CLOCKS
    if (p.dbg && blockIdx.x == 0 && (SP ? threadIdx.x == 0 : G.tl == 0) && pc[7] > 2)
        for (int i = 0; i < 8; ++i) p.dbg[i] = (int)(pc[i] >> 4);      // units of 16 cycles
#endif
    if constexpr (KS == K_TMEM) tmem_release_tables<TM_COLS>(tk);
}

// ---------------------------------------------------------------------------
// Evaluation only (unit-parity hook and the host's compute_trajectory_cost*):
// one team per trajectory, grid-stride.
// ---------------------------------------------------------------------------
template <int WPT, bool STRICT, int KS, int NW, bool ARM>
__global__ void __launch_bounds__(NW * 32) fgd_eval_kernel(const __grid_constant__ DevParams p, const EvalPtrs e)
{
    static_assert(WPT == 1 || NW == WPT, "multi-warp teams own their CTA");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TEAMS = NW / WPT;
    const int T = p.T;
    const SmemLayout L = make_layout(T, WPT * 32 * R, p.n_obs, KS, TEAMS, WPT);
    float *sKD = reinterpret_cast<float *>(smem_raw);
    float *sKO = sKD + 2 * L.k_floats;
    float2 *sObs = reinterpret_cast<float2 *>(sKO + L.k_floats);
    unsigned char *sTeams = reinterpret_cast<unsigned char *>(sObs + L.obs_pairs);
    stage_constants<KS>(p, L, sKD, sKO, sObs, NW * 32);

    const int team = (threadIdx.x >> 5) / WPT;
    float4 *XA = reinterpret_cast<float4 *>(sTeams + (size_t)team * L.team_bytes()), *XB = XA + L.x_rows;
    const Team<WPT> G(reinterpret_cast<float *>(XB + L.x_rows));
    const float *kd = (KS == K_SMEM ? sKD : p.KD) + G.tl * 2 * R;
    const int stride = gridDim.x * TEAMS;
    const int t0 = G.tl * R;

    for (int b = blockIdx.x * TEAMS + team; b < p.B; b += stride) {       // team-uniform trip count
        float start[3], goal[3];
        const float *ap = p.alpha + (size_t)b * T * 3;
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (t0 + r < T) XA[t0 + r] = make_float4(ap[(t0 + r) * 3], ap[(t0 + r) * 3 + 1], ap[(t0 + r) * 3 + 2], 0.0f);
#pragma unroll
        for (int k = 0; k < 3; ++k) { start[k] = p.start[(size_t)b * 3 + k]; goal[k] = p.goal[(size_t)b * 3 + k]; }
        f2 y1[3], y2[3];
        G.sync();
        contract<WPT, KS, true>(kd, T, XA, XA, y1, y2);
        G.sync();
        Rows<ARM> Rw;
        float loss, toc;
        int ful;
        cost_phase<WPT, STRICT, ARM>(p, T, sObs, p.n_obs, G, y1, y2, start, goal, e.lam_sg, e.lam_jl, Rw, loss, toc, ful, share_split(T, p.n_obs, ARM));
        if (G.tl == 0) {
            if (e.loss) e.loss[b] = loss;
            if (e.toc) e.toc[b] = toc;
            if (e.fulfilled) e.fulfilled[b] = ful;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (t0 < T) {
                if (e.q) e.q[((size_t)b * T + t0) * 3 + k] = Rw.q[k].x;
                if (e.v) e.v[((size_t)b * T + t0) * 3 + k] = Rw.v[k].x;
            }
            if (t0 + 1 < T) {
                if (e.q) e.q[((size_t)b * T + t0 + 1) * 3 + k] = Rw.q[k].y;
                if (e.v) e.v[((size_t)b * T + t0 + 1) * 3 + k] = Rw.v[k].y;
            }
        }
        if (e.grad) {
            unsigned nz_unused[WPT][R];
            grad_phase<WPT, ARM>(p, T, G, Rw, start, goal, e.lam_sg, e.lam_jl, XA, XB, nz_unused);
            G.sync();
            contract<WPT, KS, false>(kd, T, XA, XB, y1, y2);      // dense reference form of the backward contraction
            f2 g[3];
            backward_rows(p, y1, y2, g);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (t0 < T) e.grad[((size_t)b * T + t0) * 3 + k] = g[k].x;
                if (t0 + 1 < T) e.grad[((size_t)b * T + t0 + 1) * 3 + k] = g[k].y;
            }
        }
        G.sync();
    }
}

// ---------------------------------------------------------------------------
// Best restart per problem (random-restart sweep): one warp per problem.
// Key order: constraint-fulfilled first, then lower obstacle cost, then lower index.
// ---------------------------------------------------------------------------
__global__ void fgd_argmin_kernel(int n_problems, int n_restarts, const float *__restrict__ fstate,
                                  const int *__restrict__ istate, int index_offset, int problem_stride, float *best_cost, int *best_index,
                                  long long *best_key)
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
        const unsigned long long key = ((unsigned long long)ful << 62) | ((unsigned long long)cb << 31) | (unsigned)r;
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
        const int gi = index_offset + prob * problem_stride + r;
        if (best_cost) best_cost[prob] = fstate[b * FGD_FSTATE + FGD_F_TOC];
        if (best_index) best_index[prob] = gi;
        // same order key with the GLOBAL index in the low bits: the winner over several shards is the minimum key
        if (best_key) best_key[prob] = (long long)((best & ~0x7fffffffull) | (unsigned long long)(unsigned)gi);
    }
}

// ---------------------------------------------------------------------------
// initTrajectory (trajectory.py:73-78) as a rank-2 update: one thread per (b, t).
//   s' = start J^-1, d' = (goal - start) J^-1,  alpha[b][t][j] = fma(w[t], d'[j], u[t] * s'[j])
// HBM-bound: 12 B written per (b, t), 24 B read per trajectory.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fgd_init_kernel(int B, int T, const float *__restrict__ uw, const float *__restrict__ jinv,
                                                       const float *__restrict__ start, const float *__restrict__ goal,
                                                       float *__restrict__ alpha)
{
    const long long n = (long long)B * T;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / T), t = (int)(i - (long long)b * T);
        const float s0 = start[b * 3], s1 = start[b * 3 + 1], s2 = start[b * 3 + 2];
        const float e0 = goal[b * 3] - s0, e1 = goal[b * 3 + 1] - s1, e2 = goal[b * 3 + 2] - s2;
        const float u = uw[t], w = uw[T + t];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float sj = fmaf(s2, jinv[6 + j], fmaf(s1, jinv[3 + j], s0 * jinv[j]));
            const float dj = fmaf(e2, jinv[6 + j], fmaf(e1, jinv[3 + j], e0 * jinv[j]));
            alpha[i * 3 + j] = fmaf(w, dj, u * sj);
        }
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

// MUFU roofline probe: 8 independent chains x -> 1/x + 1 per thread (ptxas folds a bare rcp(rcp(x)) chain to x and drops
// the loop; the FADD keeps it honest and costs 1/8 of a MUFU slot on another pipe).
__global__ void __launch_bounds__(256) fgd_mufu_peak_kernel(int iters, float seed, float *sink)
{
    float a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = seed + 0.25f * u + 1e-3f * threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int u = 0; u < 8; ++u) a[u] = rcp<false>(a[u]) + 1.0f;
        }
    }
    const float r = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    if (r == 123456.789f) sink[0] = r;
}

}  // namespace fgd
