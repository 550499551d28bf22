// fgd_api.cu -- C ABI (include/fgd_b200.h) over the sm_100a kernels.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fgd_kernels.cuh"

using namespace fgd;

struct FgdHandle {
    FgdConfig cfg;
    DevParams base;            // everything except the per-call batch pointers
    int T, TP, WPT, variant = 0;
    int device, num_sms, max_smem_optin;
    float *d_KD = nullptr, *d_KO = nullptr, *d_DO = nullptr;
    float *d_init = nullptr;   // [2T + 9]: u = K^-1 1, w = K^-1 c, J^-1 (fgd_set_init_basis)
    // obstacle ring: generation g lives in slot g % FGD_OBS_RING (fgd_set_obstacles_async)
    float *d_obs_ring = nullptr;           // [FGD_OBS_RING][capacity][2]
    int *d_obs_meta = nullptr;             // [0] latest generation, [2 + 2s], [3 + 2s]: generation / count of slot s
    float *h_obs_stage = nullptr;          // page-locked staging copies of host sources, one per slot
    int *h_meta_stage = nullptr;           // page-locked sources of the header copies: per slot {-1, gen, count, gen}
    cudaEvent_t stage_event[FGD_OBS_RING] = {};
    bool stage_pending[FGD_OBS_RING] = {};
    bool slot_captured[FGD_OBS_RING] = {}; // a launch enqueued since the slot was written reads it
    int obs_gen = 0, obs_count = 0;
    unsigned *d_queue = nullptr;   // FGD_QUEUE_RING work-queue counters: launch n owns counter n % FGD_QUEUE_RING
    int *h_dbg = nullptr, *d_dbg = nullptr;   // FGD_DEBUG_MARK builds: host-mapped progress markers
    cudaEvent_t obs_event = nullptr, launch_event = nullptr;
    bool obs_event_pending = false, launch_event_pending = false;
    int spec_max_batch = 0;    // BLS batches up to this size run the speculative line-search kernel (one trajectory per CTA)
    long long spec_launches = 0;
    int last_cuda_error = 0;
    long long launches = 0, zero_copy_calls = 0, opt_launches = 0;
    // scratch for the host-buffer entry point
    float *s_alpha = nullptr, *s_start = nullptr, *s_goal = nullptr, *s_fstate = nullptr;
    int *s_istate = nullptr;
    int s_cap = 0;
};

#define CK(call)                                                     \
    do {                                                             \
        cudaError_t e_ = (call);                                     \
        if (e_ != cudaSuccess) { h->last_cuda_error = (int)e_; return FGD_ERR_CUDA; } \
    } while (0)

namespace {

struct Geometry { int grid, block, smem; };

// warps per trajectory by trajectory length (each thread owns R = 2 adjacent time samples: TP = 64 * WPT >= T)
inline int warps_per_trajectory(int T) { return T <= 64 ? 1 : (T <= 128 ? 2 : 4); }

// The smallest float t with sqrtf(t) >= eps.  The correctly rounded square root is monotonic, so for every x >= 0 (and
// NaN)  sqrtf(x) < eps  <=>  x < t : the kernels test the squared norms against t and never take a square root.
inline float sqrt_threshold(float eps)
{
    if (!(eps > 0.0f)) return 0.0f;                         // sqrtf(x) < eps is never true
    float t = eps * eps;
    while (std::sqrt(t) < eps) t = std::nextafter(t, INFINITY);
    while (t > 0.0f && std::sqrt(std::nextafter(t, 0.0f)) >= eps) t = std::nextafter(t, 0.0f);
    return t;
}

// (variant, WPT, KSRC, NW, MINB): the instantiated optimiser kernels.  WPT warps per trajectory; KSRC: where the K tables
// live - tensor memory (K_TMEM), shared memory (K_SMEM, T <= 64) or L2 (K_L2); NW warps per CTA, MINB = min CTAs per SM
// (register cap).  Variant 0 is the default: ONE 16-warp CTA per SM with the tables in tensor memory - 16 single-warp teams
// (T <= 64), 8 two-warp teams (T <= 128, K and dK in TMEM) or 4 four-warp teams (T <= 256, K in TMEM, dK from L2), the
// multi-warp teams on named barriers.  Variant 1 exists for A/B measurements (env FGD_VARIANT): shared-memory tables for
// T <= 64, one team per CTA with the tables in L2 for T > 64 (the round-1 layout).
#define FGD_FOR_CONFIGS(X) \
    X(0, 1, K_TMEM, 16, 1) X(0, 2, K_TMEM, 16, 1) X(0, 4, K_TMEM, 16, 1) \
    X(1, 1, K_SMEM, 8, 2) X(1, 2, K_L2, 2, 8) X(1, 4, K_L2, 4, 4)
// (WPT, KSRC, NW): the evaluation kernels (parity hook; tables in shared memory / L2)
#define FGD_FOR_EVAL_CONFIGS(X) X(1, K_SMEM, 8) X(2, K_L2, 2) X(4, K_L2, 4)

// The T = 50 instances have a twin for the default scene's obstacle count (FGD_NO_OC=1 disables it for A/B measurements).
constexpr int FGD_OC = 11;
inline bool use_oc()
{
    static const bool off = [] { const char *e = std::getenv("FGD_NO_OC"); return e && e[0] == '1'; }();
    return !off;
}

template <int WPT, bool STRICT, int KS, int NW, int MINB, bool ARM, int TC>
cudaError_t launch_opt(const DevParams &p, int grid, size_t smem, cudaStream_t st)
{
    auto kern = p.mode == 0 ? fgd_optimize_kernel<WPT, STRICT, KS, NW, MINB, ARM, TC, 0> : fgd_optimize_kernel<WPT, STRICT, KS, NW, MINB, ARM, TC, 1>;
    if constexpr (WPT == 1 && KS == K_TMEM && !ARM) {
        // many obstacles: the instance whose sample-less lanes take a share of the obstacle loop (same bits, same launch bounds)
        if (share_split(p.T, p.n_obs, false) > 0)
            kern = p.mode == 0 ? fgd_optimize_kernel<WPT, STRICT, KS, NW, MINB, ARM, TC, 0, 0, false, true>
                               : fgd_optimize_kernel<WPT, STRICT, KS, NW, MINB, ARM, TC, 1, 0, false, true>;
        if constexpr (TC > 0) {
            // the reference's default scene (environment.py:17-29: 11 obstacles) next to its default T: unrolled obstacle loop
            if (p.n_obs == FGD_OC && use_oc())
                kern = p.mode == 0 ? fgd_optimize_kernel<WPT, STRICT, KS, NW, MINB, ARM, TC, 0, 0, false, false, FGD_OC>
                                   : fgd_optimize_kernel<WPT, STRICT, KS, NW, MINB, ARM, TC, 1, 0, false, false, FGD_OC>;
        }
    }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, NW * 32, smem, st>>>(p);
    return cudaGetLastError();
}

template <int WPT, bool STRICT, int KS, int NW, int MINB, bool ARM, int TC>
int occupancy_opt(size_t smem)
{
    auto kern = fgd_optimize_kernel<WPT, STRICT, KS, NW, MINB, ARM, TC, 0>;      // both modes share the launch bounds
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, NW * 32, smem) != cudaSuccess) nb = 1;
    return nb < 1 ? 1 : nb;
}

template <int WPT, bool STRICT, int KS, int NW, bool ARM>
cudaError_t launch_eval(const DevParams &p, const EvalPtrs &e, int grid, size_t smem, cudaStream_t st)
{
    auto kern = fgd_eval_kernel<WPT, STRICT, KS, NW, ARM>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    kern<<<grid, NW * 32, smem, st>>>(p, e);
    return cudaGetLastError();
}

// Speculative line search (latency mode, fgd_kernels.cuh): SPEC_WARPS replicas of one trajectory per CTA, tables in TMEM.
constexpr int SPEC_WARPS = 4;
template <bool STRICT, int TC>
cudaError_t launch_spec(const DevParams &p, int grid, size_t smem, cudaStream_t st)
{
    auto kern = fgd_optimize_kernel<1, STRICT, K_TMEM, SPEC_WARPS, 2, false, TC, 0, SPEC_WARPS>;
    if (share_split(p.T, p.n_obs, false) > 0)          // many obstacles: the replicas' sample-less lanes share the obstacle loop
        kern = fgd_optimize_kernel<1, STRICT, K_TMEM, SPEC_WARPS, 2, false, TC, 0, SPEC_WARPS, false, true>;
    if constexpr (TC > 0) {
        if (p.n_obs == FGD_OC && use_oc())              // the default scene: unrolled obstacle loop
            kern = fgd_optimize_kernel<1, STRICT, K_TMEM, SPEC_WARPS, 2, false, TC, 0, SPEC_WARPS, false, false, FGD_OC>;
    }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, SPEC_WARPS * 32, smem, st>>>(p);
    return cudaGetLastError();
}

// Live obstacle updates (fgd_optimize_live): the default T <= 64 kernel layout with a private obstacle set per team.
template <bool STRICT, int TC>
cudaError_t launch_live(const DevParams &p, int grid, size_t smem, cudaStream_t st)
{
    auto kern = p.mode == 0 ? fgd_optimize_kernel<1, STRICT, K_TMEM, 16, 1, false, TC, 0, 0, true> : fgd_optimize_kernel<1, STRICT, K_TMEM, 16, 1, false, TC, 1, 0, true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, 16 * 32, smem, st>>>(p);
    return cudaGetLastError();
}

bool variant_exists(int v, int WPT)
{
#define X(V_, W_, KS_, NW_, MB_) if (v == V_ && WPT == W_) return true;
    FGD_FOR_CONFIGS(X)
#undef X
    return false;
}

int k_source(int v, int WPT)
{
#define X(V_, W_, KS_, NW_, MB_) if (v == V_ && WPT == W_) return KS_;
    FGD_FOR_CONFIGS(X)
#undef X
    return K_L2;
}

int eval_k_source(int WPT)
{
#define X(W_, KS_, NW_) if (WPT == W_) return KS_;
    FGD_FOR_EVAL_CONFIGS(X)
#undef X
    return K_L2;
}

int eval_warps_per_cta(int WPT)
{
#define X(W_, KS_, NW_) if (WPT == W_) return NW_;
    FGD_FOR_EVAL_CONFIGS(X)
#undef X
    return WPT;
}

int warps_per_cta(int v, int WPT)
{
#define X(V_, W_, KS_, NW_, MB_) if (v == V_ && WPT == W_) return NW_;
    FGD_FOR_CONFIGS(X)
#undef X
    return WPT;
}

// The tensor-memory kernels have an instance specialised for the reference's default T (main.py: --n-timesteps 50).
constexpr int FGD_TC = 50;
inline bool use_tc(int T)
{
    static const bool off = [] { const char *e = std::getenv("FGD_NO_TC"); return e && e[0] == '1'; }();
    return !off && T == FGD_TC;
}

// the whole-arm cost variants exist for runtime T only (TC = 0)
template <int W, int KS, int NW, int MB, int TC>
cudaError_t launch_sa(bool strict, bool arm, const DevParams &p, int grid, size_t smem, cudaStream_t st)
{
    return arm ? (strict ? launch_opt<W, true, KS, NW, MB, true, 0>(p, grid, smem, st) : launch_opt<W, false, KS, NW, MB, true, 0>(p, grid, smem, st))
               : (strict ? launch_opt<W, true, KS, NW, MB, false, TC>(p, grid, smem, st) : launch_opt<W, false, KS, NW, MB, false, TC>(p, grid, smem, st));
}

template <int W, int KS, int NW, int MB, int TC>
int occupancy_sa(bool strict, bool arm, size_t smem)
{
    return arm ? (strict ? occupancy_opt<W, true, KS, NW, MB, true, 0>(smem) : occupancy_opt<W, false, KS, NW, MB, true, 0>(smem))
               : (strict ? occupancy_opt<W, true, KS, NW, MB, false, TC>(smem) : occupancy_opt<W, false, KS, NW, MB, false, TC>(smem));
}

cudaError_t dispatch_opt(int v, int WPT, bool strict, bool arm, const DevParams &p, int grid, size_t smem, cudaStream_t st)
{
#define X(V_, W_, KS_, NW_, MB_)                                                                             \
    if (v == V_ && WPT == W_) {                                                                               \
        if constexpr (KS_ == K_TMEM && W_ == 1) {                                                             \
            if (use_tc(p.T)) return launch_sa<W_, KS_, NW_, MB_, FGD_TC>(strict, arm, p, grid, smem, st);     \
        }                                                                                                     \
        return launch_sa<W_, KS_, NW_, MB_, 0>(strict, arm, p, grid, smem, st);                               \
    }
    FGD_FOR_CONFIGS(X)
#undef X
    return cudaErrorInvalidValue;
}

int dispatch_occ(int v, int WPT, int T, bool strict, bool arm, size_t smem)
{
#define X(V_, W_, KS_, NW_, MB_)                                                                             \
    if (v == V_ && WPT == W_) {                                                                               \
        if constexpr (KS_ == K_TMEM && W_ == 1) {                                                             \
            if (use_tc(T)) return occupancy_sa<W_, KS_, NW_, MB_, FGD_TC>(strict, arm, smem);                 \
        }                                                                                                     \
        return occupancy_sa<W_, KS_, NW_, MB_, 0>(strict, arm, smem);                                         \
    }
    FGD_FOR_CONFIGS(X)
#undef X
    return 1;
}

cudaError_t dispatch_eval(int WPT, bool strict, bool arm, const DevParams &p, const EvalPtrs &e, int grid, size_t smem, cudaStream_t st)
{
#define X(W_, KS_, NW_)                                                                 \
    if (WPT == W_)                                                                       \
        return arm ? (strict ? launch_eval<W_, true, KS_, NW_, true>(p, e, grid, smem, st)       \
                             : launch_eval<W_, false, KS_, NW_, true>(p, e, grid, smem, st))     \
                   : (strict ? launch_eval<W_, true, KS_, NW_, false>(p, e, grid, smem, st)      \
                             : launch_eval<W_, false, KS_, NW_, false>(p, e, grid, smem, st));
    FGD_FOR_EVAL_CONFIGS(X)
#undef X
    return cudaErrorInvalidValue;
}

Geometry geometry(const FgdHandle *h, int B, int n_obs)
{
    Geometry g;
    const int nw = warps_per_cta(h->variant, h->WPT);
    const int teams = nw / h->WPT;
    g.block = nw * 32;
    g.smem = (int)make_layout(h->T, h->TP, n_obs, k_source(h->variant, h->WPT), teams, h->WPT).bytes();
    const int occ = dispatch_occ(h->variant, h->WPT, h->T, h->cfg.strict_math != 0, h->cfg.whole_arm_cost != 0, (size_t)g.smem);
    // one CTA per trajectory until every SM has its share: a batch smaller than the machine spreads over all SMs (the
    // first pick of every team is static, fetch_slot), larger batches run occupancy x SMs persistent CTAs
    const long long need = (long long)B;
    const long long cap = (long long)occ * h->num_sms;
    g.grid = (int)(need < cap ? need : cap);
    if (g.grid < 1) g.grid = 1;
    return g;
}

void fill_params(const FgdHandle *h, DevParams &p, int mode, int B, float *alpha, const float *start, const float *goal,
                 float *fstate, int *istate, int budget)
{
    p = h->base;
    p.mode = mode; p.B = B; p.budget = budget;
    p.n_obs = h->obs_count;
    p.obs = h->d_obs_ring + (size_t)(h->obs_gen % FGD_OBS_RING) * h->cfg.obstacle_capacity * 2;
    p.obs_meta = h->d_obs_meta; p.obs_ring = h->d_obs_ring; p.obs_cap = h->cfg.obstacle_capacity;
    p.poll_every = 0; p.switch_log = nullptr;
    p.alpha = alpha; p.alpha_in = alpha; p.fresh = 0;
    p.start = start; p.goal = goal; p.fstate = fstate; p.istate = istate;
    p.queue = h->d_queue + (h->opt_launches % FGD_QUEUE_RING);
    p.dbg = h->d_dbg;
}

int wait_obstacles(FgdHandle *h, cudaStream_t st)
{
    if (h->obs_event_pending) {
        CK(cudaStreamWaitEvent(st, h->obs_event, 0));
    }
    return FGD_OK;
}

// one persistent launch of the optimiser (p is complete; its queue counter has been reset on `st`)
int launch_one(FgdHandle *h, const DevParams &p, int mode, int B, bool live, bool spec, cudaStream_t st)
{
    Geometry g = geometry(h, B, p.n_obs);
    const bool strict = h->cfg.strict_math != 0;
    if (live) {         // always the tensor-memory layout: 16 single-warp teams per CTA, one CTA per SM
        g.block = 16 * 32;
        g.smem = (int)make_layout(h->T, h->TP, h->cfg.obstacle_capacity, K_TMEM, 16, 1, true).bytes();
        g.grid = B < h->num_sms ? B : h->num_sms;
    }
    if (g.smem + 64 > h->max_smem_optin) return FGD_ERR_TOO_MANY_OBSTACLES;
    if (spec) {
        g.block = SPEC_WARPS * 32;
        g.smem = (int)make_layout(h->T, h->TP, p.n_obs, K_TMEM, SPEC_WARPS, 1).bytes();
        g.grid = B < 2 * h->num_sms ? B : 2 * h->num_sms;
        CK(use_tc(p.T) ? (strict ? launch_spec<true, FGD_TC>(p, g.grid, (size_t)g.smem, st) : launch_spec<false, FGD_TC>(p, g.grid, (size_t)g.smem, st))
                       : (strict ? launch_spec<true, 0>(p, g.grid, (size_t)g.smem, st) : launch_spec<false, 0>(p, g.grid, (size_t)g.smem, st)));
        h->spec_launches += 1;
    } else if (live) {
        CK(use_tc(p.T) ? (strict ? launch_live<true, FGD_TC>(p, g.grid, (size_t)g.smem, st) : launch_live<false, FGD_TC>(p, g.grid, (size_t)g.smem, st))
                       : (strict ? launch_live<true, 0>(p, g.grid, (size_t)g.smem, st) : launch_live<false, 0>(p, g.grid, (size_t)g.smem, st)));
    } else
        CK(dispatch_opt(h->variant, h->WPT, strict, h->cfg.whole_arm_cost != 0, p, g.grid, (size_t)g.smem, st));
    h->launches += 1;
    h->opt_launches += 1;
    if (!live) {        // a live kernel copies what it needs and validates it against the slot header; it must not hold up the publisher
        CK(cudaEventRecord(h->launch_event, st));
        h->launch_event_pending = true;
        h->slot_captured[h->obs_gen % FGD_OBS_RING] = true;
    }
    return FGD_OK;
}

// d_alpha_in != nullptr: a fresh run that reads the initial alpha rows from d_alpha_in, writes the results to d_alpha
// and never reads the loop state (fgd_optimize_host_io).
int run_optimize(FgdHandle *h, int mode, int B, float *d_alpha, const float *d_start, const float *d_goal, float *d_fstate,
                 int *d_istate, int budget, cudaStream_t st, const float *d_alpha_in = nullptr, int poll_every = 0, int *d_switch_log = nullptr)
{
    if (!h || B < 0 || (B > 0 && (!d_alpha || !d_start || !d_goal || !d_fstate || !d_istate))) return FGD_ERR_INVALID_ARGUMENT;
    if (mode == 1 && h->cfg.max_outer_iteration > h->cfg.n_gd_lr) return FGD_ERR_INVALID_ARGUMENT;   // optimizer_GD.py:34-36
    if (B == 0) return FGD_OK;
    int rc = wait_obstacles(h, st);
    if (rc) return rc;
    DevParams p;
    fill_params(h, p, mode, B, d_alpha, d_start, d_goal, d_fstate, d_istate, budget);
    if (d_alpha_in) { p.alpha_in = d_alpha_in; p.fresh = 1; }
    const bool live = poll_every > 0;
    if (live) {
        if (h->WPT != 1) return FGD_ERR_UNSUPPORTED_T;
        if (h->cfg.whole_arm_cost) return FGD_ERR_INVALID_ARGUMENT;
        p.poll_every = poll_every; p.switch_log = d_switch_log;
    }
    const bool spec = mode == 0 && h->WPT == 1 && h->variant == 0 && !h->cfg.whole_arm_cost && !live && B <= h->spec_max_batch;
    CK(cudaMemsetAsync(p.queue, 0, sizeof(unsigned), st));      // this launch's own counter (launches on other streams keep theirs)
    return launch_one(h, p, mode, B, live, spec, st);
}

// Page-locked host memory is addressable from the device (unified addressing): returns the device alias of
// [p, p + bytes) if the WHOLE range is page-locked and mapped contiguously (first and last byte are checked: a buffer
// whose registration covers only part of the range must take the staged path, not fault on the device).
bool mapped_host_range(const void *p, size_t bytes, void **dev)
{
    if (bytes == 0) return false;
    cudaPointerAttributes a0, a1;
    if (cudaPointerGetAttributes(&a0, p) != cudaSuccess) { cudaGetLastError(); return false; }
    if (a0.type != cudaMemoryTypeHost || !a0.devicePointer) return false;
    const char *last = static_cast<const char *>(p) + (bytes - 1);
    if (cudaPointerGetAttributes(&a1, last) != cudaSuccess) { cudaGetLastError(); return false; }
    if (a1.type != cudaMemoryTypeHost || !a1.devicePointer) return false;
    if (static_cast<const char *>(a1.devicePointer) - static_cast<const char *>(a0.devicePointer) != (ptrdiff_t)(bytes - 1)) return false;
    *dev = a0.devicePointer;
    return true;
}

int ensure_scratch(FgdHandle *h, int B)
{
    if (B > h->s_cap) {
        cudaFree(h->s_alpha); cudaFree(h->s_start); cudaFree(h->s_goal); cudaFree(h->s_fstate); cudaFree(h->s_istate);
        h->s_alpha = h->s_start = h->s_goal = h->s_fstate = nullptr; h->s_istate = nullptr; h->s_cap = 0;
        CK(cudaMalloc(&h->s_alpha, (size_t)B * h->T * 3 * 4));
        CK(cudaMalloc(&h->s_start, (size_t)B * 3 * 4));
        CK(cudaMalloc(&h->s_goal, (size_t)B * 3 * 4));
        CK(cudaMalloc(&h->s_fstate, (size_t)B * FGD_FSTATE * 4));
        CK(cudaMalloc(&h->s_istate, (size_t)B * FGD_ISTATE * 4));
        h->s_cap = B;
    }
    return FGD_OK;
}

// best-of-5 throughput of a probe kernel; all resources are released on every path
template <typename Launch>
int measure_peak(FgdHandle *h, cudaStream_t st, double work_per_launch, Launch launch, double *out)
{
    float *sink = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    double best = 0.0;
    cudaError_t err = cudaMalloc(&sink, 4);
    if (err == cudaSuccess) err = cudaEventCreate(&e0);
    if (err == cudaSuccess) err = cudaEventCreate(&e1);
    for (int rep = 0; rep < 6 && err == cudaSuccess; ++rep) {          // first rep warms up, best of the rest
        err = cudaEventRecord(e0, st);
        if (err != cudaSuccess) break;
        launch(1.0f + rep, sink);
        err = cudaGetLastError();
        if (err == cudaSuccess) err = cudaEventRecord(e1, st);
        if (err == cudaSuccess) err = cudaEventSynchronize(e1);
        float ms = 0.f;
        if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e0, e1);
        if (err == cudaSuccess && rep > 0 && ms > 0.f) { const double t = work_per_launch / (ms * 1e-3) * 1e-12; if (t > best) best = t; }
        h->launches += 1;
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(sink);
    if (err != cudaSuccess) { h->last_cuda_error = (int)err; return FGD_ERR_CUDA; }
    *out = best;
    return FGD_OK;
}

}  // namespace

extern "C" {

int fgd_abi_version(void) { return FGD_ABI_VERSION; }

float fgd_sqrt_threshold(float eps) { return sqrt_threshold(eps); }

const char *fgd_status_string(int s)
{
    switch (s) {
        case FGD_OK: return "ok";
        case FGD_ERR_INVALID_ARGUMENT: return "invalid argument";
        case FGD_ERR_UNSUPPORTED_T: return "n_timesteps outside [2, 256]";
        case FGD_ERR_KERNEL_NOT_SYMMETRIC: return "km must be symmetric and dkm antisymmetric (bit-wise)";
        case FGD_ERR_TOO_MANY_OBSTACLES: return "obstacle count exceeds obstacle_capacity (or the capacity does not fit in shared memory)";
        case FGD_ERR_CUDA: return "CUDA error (see fgd_last_cuda_error)";
        case FGD_ERR_NO_DEVICE: return "no CUDA device";
        case FGD_ERR_JOINTS: return "n_joints must be 3";
        default: return "unknown status";
    }
}

int fgd_last_cuda_error(const FgdHandle *h) { return h ? h->last_cuda_error : 0; }

int fgd_create(const FgdConfig *cfg, FgdHandle **out)
{
    if (!cfg || !out || cfg->abi_version != FGD_ABI_VERSION || !cfg->h_km || !cfg->h_dkm) return FGD_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (cfg->n_joints != 3) return FGD_ERR_JOINTS;
    const int T = cfg->n_timesteps;
    if (T < 2 || T > FGD_MAX_T) return FGD_ERR_UNSUPPORTED_T;
    if (cfg->obstacle_capacity < 1 || cfg->max_outer_iteration < 1 || cfg->max_outer_iteration > FGD_MAX_OUTER ||
        cfg->n_gd_lr < 0 || cfg->n_gd_lr > FGD_MAX_OUTER || cfg->max_bls_iteration < 1)
        return FGD_ERR_INVALID_ARGUMENT;
    // K^T = K and dK^T = -dK bit-wise: lets the backward contraction reuse the staged tiles.
    for (int i = 0; i < T; ++i)
        for (int k = 0; k < T; ++k) {
            const float a = cfg->h_km[i * T + k], b = cfg->h_km[k * T + i];
            const float c = cfg->h_dkm[i * T + k], d = -cfg->h_dkm[k * T + i];
            if (!(a == b)) return FGD_ERR_KERNEL_NOT_SYMMETRIC;
            if (!(c == d)) return FGD_ERR_KERNEL_NOT_SYMMETRIC;
        }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return FGD_ERR_NO_DEVICE;

    FgdHandle *h = new FgdHandle();
    h->cfg = *cfg;
    h->cfg.h_km = nullptr; h->cfg.h_dkm = nullptr;
    h->T = T;
    h->WPT = warps_per_trajectory(T);
    h->TP = h->WPT * 32 * R;
    auto fail = [&](int code) { fgd_destroy(h); return code; };
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(FGD_ERR_CUDA); } while (0)
    CKC(cudaGetDevice(&h->device));
    CKC(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device));
    CKC(cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    // K and dK go to shared memory when both fit beside the operand buffers of at least one slot per warp
    if (const char *e = std::getenv("FGD_VARIANT")) { const int v = std::atoi(e); if (variant_exists(v, h->WPT)) h->variant = v; }
    h->spec_max_batch = 2 * h->num_sms;          // up to two 4-warp CTAs per SM: beyond that the warps are better used on more trajectories
    if (const char *e = std::getenv("FGD_SPEC_MAX_BATCH")) h->spec_max_batch = std::atoi(e);      // 0 disables (A/B measurements)
    {   // the whole obstacle set is staged in shared memory next to the operand buffers (and, for the kernels that keep
        // them there, the K tables): the capacity must fit both the optimiser and the evaluation kernel
        const int nw = warps_per_cta(h->variant, h->WPT), nwe = eval_warps_per_cta(h->WPT);
        const size_t need = make_layout(T, h->TP, cfg->obstacle_capacity, k_source(h->variant, h->WPT), nw / h->WPT, h->WPT).bytes() + 64;
        const size_t need_e = make_layout(T, h->TP, cfg->obstacle_capacity, eval_k_source(h->WPT), nwe / h->WPT, h->WPT).bytes() + 64;
        if (need > (size_t)h->max_smem_optin || need_e > (size_t)h->max_smem_optin) return fail(FGD_ERR_TOO_MANY_OBSTACLES);
    }

    // operand table KD[k][thread][2R]: the R row entries K[t][k] then the R entries dK[t][k] of the team thread's rows
    // t = R*thread + r, and KO[k][thread][R]: the K entries alone (dense half of the backward contraction)
    const int nthr = h->TP / R;
    std::vector<float> kd((size_t)T * 2 * h->TP, 0.0f), ko((size_t)T * h->TP, 0.0f), dko((size_t)T * h->TP, 0.0f);
    for (int k = 0; k < T; ++k)
        for (int i = 0; i < T; ++i) {
            const int thr = i / R, r = i % R;
            const size_t base = ((size_t)k * nthr + thr) * 2 * R;
            kd[base + r] = cfg->h_km[i * T + k];
            kd[base + R + r] = cfg->h_dkm[i * T + k];
            ko[((size_t)k * nthr + thr) * R + r] = cfg->h_km[i * T + k];
            dko[((size_t)k * nthr + thr) * R + r] = cfg->h_dkm[i * T + k];
        }
    CKC(cudaMalloc(&h->d_KD, kd.size() * 4));
    CKC(cudaMemcpy(h->d_KD, kd.data(), kd.size() * 4, cudaMemcpyHostToDevice));
    CKC(cudaMalloc(&h->d_KO, ko.size() * 4));
    CKC(cudaMemcpy(h->d_KO, ko.data(), ko.size() * 4, cudaMemcpyHostToDevice));
    CKC(cudaMalloc(&h->d_DO, dko.size() * 4));
    CKC(cudaMemcpy(h->d_DO, dko.data(), dko.size() * 4, cudaMemcpyHostToDevice));
    {
        const size_t slot = (size_t)cfg->obstacle_capacity * 2 * 4;
        CKC(cudaMalloc(&h->d_obs_ring, FGD_OBS_RING * slot));
        CKC(cudaMemset(h->d_obs_ring, 0, FGD_OBS_RING * slot));
        std::vector<int> meta(2 + 2 * FGD_OBS_RING, -1);       // generation 0 = the empty set in slot 0
        meta[0] = 0; meta[1] = 0; meta[2] = 0; meta[3] = 0;
        CKC(cudaMalloc(&h->d_obs_meta, meta.size() * 4));
        CKC(cudaMemcpy(h->d_obs_meta, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice));
        CKC(cudaHostAlloc(&h->h_obs_stage, FGD_OBS_RING * slot, cudaHostAllocDefault));
        CKC(cudaHostAlloc(&h->h_meta_stage, FGD_OBS_RING * 4 * sizeof(int), cudaHostAllocDefault));
        for (int i = 0; i < FGD_OBS_RING; ++i) CKC(cudaEventCreateWithFlags(&h->stage_event[i], cudaEventDisableTiming));
    }
    CKC(cudaMalloc(&h->d_queue, FGD_QUEUE_RING * sizeof(unsigned)));
#if defined(FGD_DEBUG_MARK) || defined(FGD_PHASE_CLOCKS)
    CKC(cudaHostAlloc(&h->h_dbg, 4096 * sizeof(int), cudaHostAllocMapped));
    std::memset(h->h_dbg, 0, 4096 * sizeof(int));
    CKC(cudaHostGetDevicePointer(&h->d_dbg, h->h_dbg, 0));
#endif
    CKC(cudaEventCreateWithFlags(&h->obs_event, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&h->launch_event, cudaEventDisableTiming));
#undef CKC

    DevParams &p = h->base;
    std::memset(&p, 0, sizeof p);
    p.T = T; p.TP = h->TP;
    p.max_inner = cfg->max_inner_iteration; p.max_outer = cfg->max_outer_iteration; p.max_bls = cfg->max_bls_iteration;
    p.cvdl = cfg->constraint_violating_dependant_loss ? 1 : 0;
    p.lam_sg0 = cfg->lambda_sg_constraint; p.lam_jl0 = cfg->lambda_jl_constraint; p.lam_inc = cfg->lambda_constraint_increase;
    p.lam_max = cfg->lambda_max_cost; p.lam_reg = cfg->lambda_reg; p.eps_loop = cfg->loop_loss_reduction;
    p.eps_pos = cfg->eps_position; p.eps_vel = cfg->eps_velocity;
    p.thr_pos = sqrt_threshold(p.eps_pos); p.thr_vel = sqrt_threshold(p.eps_vel);
    p.bls_lr0 = cfg->bls_lr_start; p.bls_alpha = cfg->bls_alpha; p.bls_bp = cfg->bls_beta_plus; p.bls_bm = cfg->bls_beta_minus;
    p.qmax = cfg->max_joint_position; p.qmin = cfg->min_joint_position; p.vmax = cfg->max_joint_velocity;
    for (int i = 0; i < 3; ++i) p.link[i] = cfg->link_length[i];
    for (int i = 0; i < 9; ++i) p.J[i] = cfg->jac[i];
    for (int i = 0; i < 16; ++i) p.gd_lr[i] = i < cfg->n_gd_lr ? cfg->gd_lr[i] : (cfg->n_gd_lr > 0 ? cfg->gd_lr[cfg->n_gd_lr - 1] : 0.0f);
    // derived constants: the same FP32 expressions as oracle/fgd_mirror.c derive()
    p.fT = (float)T;
    p.oml = 1.0f - p.lam_max;
    p.inv_T = 1.0f / p.fT;
    p.w_avg = p.oml * p.inv_T;
    p.mean_q = 0.5f * (p.qmax + p.qmin);                        // trajectory.py:31
    const float std_q = 0.5f * (p.qmax - p.mean_q);             // trajectory.py:32
    p.inv_std = 1.0f / std_q;
    p.inv_std2 = 1.0f / (std_q * std_q);
    p.inv_vmax = 1.0f / p.vmax;
    p.inv_vmax2 = 1.0f / (p.vmax * p.vmax);
    p.q_hi = cfg->joint_safety_limit * p.qmax;
    p.q_lo = cfg->joint_safety_limit * p.qmin;
    p.v_hi = cfg->joint_safety_limit * p.vmax;
    p.KD = h->d_KD; p.KO = h->d_KO; p.DO = h->d_DO;
    *out = h;
    return FGD_OK;
}

int fgd_destroy(FgdHandle *h)
{
    if (!h) return FGD_OK;
    cudaFree(h->d_init);
    cudaFree(h->d_KD); cudaFree(h->d_KO); cudaFree(h->d_DO); cudaFree(h->d_obs_ring); cudaFree(h->d_obs_meta); cudaFree(h->d_queue);
    if (h->h_obs_stage) cudaFreeHost(h->h_obs_stage);
    if (h->h_meta_stage) cudaFreeHost(h->h_meta_stage);
    for (int i = 0; i < FGD_OBS_RING; ++i) if (h->stage_event[i]) cudaEventDestroy(h->stage_event[i]);
    cudaFree(h->s_alpha); cudaFree(h->s_start); cudaFree(h->s_goal); cudaFree(h->s_fstate); cudaFree(h->s_istate);
    if (h->obs_event) cudaEventDestroy(h->obs_event);
    if (h->launch_event) cudaEventDestroy(h->launch_event);
    delete h;
    return FGD_OK;
}

int fgd_set_obstacles_async(FgdHandle *h, const float *xy, int32_t count, int32_t on_device, void *stream)
{
    if (!h || count < 0 || (count > 0 && !xy)) return FGD_ERR_INVALID_ARGUMENT;
    if (count > h->cfg.obstacle_capacity) return FGD_ERR_TOO_MANY_OBSTACLES;
    cudaStream_t st = (cudaStream_t)stream;
    const int gen = h->obs_gen + 1, s = gen % FGD_OBS_RING;
    const size_t slot_floats = (size_t)h->cfg.obstacle_capacity * 2;
    // the staging copies of this slot were last used FGD_OBS_RING generations ago
    if (h->stage_pending[s]) { CK(cudaEventSynchronize(h->stage_event[s])); h->stage_pending[s] = false; }
    // a launch that captured this slot may not have run yet (live kernels validate against the slot header instead)
    if (h->slot_captured[s] && h->launch_event_pending) CK(cudaStreamWaitEvent(st, h->launch_event, 0));
    h->slot_captured[s] = false;
    int *ms = h->h_meta_stage + 4 * s;
    ms[0] = -1; ms[1] = gen; ms[2] = count; ms[3] = gen;
    int *meta = h->d_obs_meta;
    float *dst = h->d_obs_ring + (size_t)s * slot_floats;
    CK(cudaMemcpyAsync(meta + 2 + 2 * s, ms, 4, cudaMemcpyHostToDevice, st));                    // header := invalid
    if (count > 0) {
        if (on_device) {
            CK(cudaMemcpyAsync(dst, xy, (size_t)count * 8, cudaMemcpyDeviceToDevice, st));
        } else {
            float *stage = h->h_obs_stage + (size_t)s * slot_floats;
            std::memcpy(stage, xy, (size_t)count * 8);                                          // the caller's buffer is free again
            CK(cudaMemcpyAsync(dst, stage, (size_t)count * 8, cudaMemcpyHostToDevice, st));
        }
    }
    CK(cudaMemcpyAsync(meta + 2 + 2 * s, ms + 1, 8, cudaMemcpyHostToDevice, st));                // header := (gen, count)
    CK(cudaMemcpyAsync(meta, ms + 3, 4, cudaMemcpyHostToDevice, st));                            // latest := gen
    CK(cudaEventRecord(h->stage_event[s], st));
    h->stage_pending[s] = true;
    CK(cudaEventRecord(h->obs_event, st));
    h->obs_event_pending = true;
    h->obs_gen = gen;
    h->obs_count = count;
    return FGD_OK;
}

int fgd_obstacle_count(const FgdHandle *h) { return h ? h->obs_count : -1; }
int fgd_obstacle_generation(const FgdHandle *h) { return h ? h->obs_gen : -1; }

int fgd_eval_cost_grad(FgdHandle *h, int32_t B, const float *d_alpha, const float *d_start, const float *d_goal,
                       float lambda_sg, float lambda_jl, float lambda_max_cost, float *d_loss, float *d_toc, float *d_grad,
                       float *d_q, float *d_v, int32_t *d_fulfilled, void *stream)
{
    if (!h || B < 0 || (B > 0 && (!d_alpha || !d_start || !d_goal))) return FGD_ERR_INVALID_ARGUMENT;
    if (B == 0) return FGD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = wait_obstacles(h, st);
    if (rc) return rc;
    DevParams p;
    fill_params(h, p, 0, B, const_cast<float *>(d_alpha), d_start, d_goal, nullptr, nullptr, -1);
    if (lambda_max_cost >= 0.0f) {      // static argument of compute_trajectory_cost (trajectory.py:271)
        p.lam_max = lambda_max_cost; p.oml = 1.0f - p.lam_max; p.w_avg = p.oml * p.inv_T;
    }
    EvalPtrs e{lambda_sg, lambda_jl, d_loss, d_toc, d_grad, d_q, d_v, d_fulfilled};
    const int nw = eval_warps_per_cta(h->WPT), per_cta = nw / h->WPT;
    const size_t smem = make_layout(h->T, h->TP, p.n_obs, eval_k_source(h->WPT), per_cta, h->WPT).bytes();
    long long need = ((long long)B + per_cta - 1) / per_cta, cap = (long long)h->num_sms * 4;
    const int grid = (int)(need < cap ? need : cap);
    CK(dispatch_eval(h->WPT, h->cfg.strict_math != 0, h->cfg.whole_arm_cost != 0, p, e, grid, smem, st));
    h->launches += 1;
    CK(cudaEventRecord(h->launch_event, st));      // this kernel reads the obstacle ring too
    h->launch_event_pending = true;
    h->slot_captured[h->obs_gen % FGD_OBS_RING] = true;
    return FGD_OK;
}

int fgd_optimize_bls(FgdHandle *h, int32_t B, float *d_alpha, const float *d_start, const float *d_goal, float *d_fstate,
                     int32_t *d_istate, int32_t max_launch_iters, void *stream)
{
    return run_optimize(h, 0, B, d_alpha, d_start, d_goal, d_fstate, d_istate, max_launch_iters, (cudaStream_t)stream);
}

int fgd_optimize_gd(FgdHandle *h, int32_t B, float *d_alpha, const float *d_start, const float *d_goal, float *d_fstate,
                    int32_t *d_istate, int32_t max_launch_iters, void *stream)
{
    return run_optimize(h, 1, B, d_alpha, d_start, d_goal, d_fstate, d_istate, max_launch_iters, (cudaStream_t)stream);
}

int fgd_optimize_live(FgdHandle *h, int32_t use_gd, int32_t B, float *d_alpha, const float *d_start, const float *d_goal,
                      float *d_fstate, int32_t *d_istate, int32_t poll_every, int32_t *d_switch_log, void *stream)
{
    if (poll_every < 1) return FGD_ERR_INVALID_ARGUMENT;
    return run_optimize(h, use_gd ? 1 : 0, B, d_alpha, d_start, d_goal, d_fstate, d_istate, -1, (cudaStream_t)stream, nullptr, poll_every,
                        d_switch_log);
}

int fgd_optimize_host(FgdHandle *h, int32_t use_gd, int32_t B, float *h_alpha, const float *h_start, const float *h_goal,
                      float *h_fstate, int32_t *h_istate, void *stream)
{
    if (!h || B < 0 || (B > 0 && (!h_alpha || !h_start || !h_goal || !h_fstate || !h_istate))) return FGD_ERR_INVALID_ARGUMENT;
    if (B == 0) return FGD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    {
        int rc0 = ensure_scratch(h, B);
        if (rc0) return rc0;
    }
    CK(cudaMemcpyAsync(h->s_alpha, h_alpha, (size_t)B * h->T * 3 * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->s_start, h_start, (size_t)B * 3 * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->s_goal, h_goal, (size_t)B * 3 * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->s_fstate, h_fstate, (size_t)B * FGD_FSTATE * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->s_istate, h_istate, (size_t)B * FGD_ISTATE * 4, cudaMemcpyHostToDevice, st));
    int rc = run_optimize(h, use_gd ? 1 : 0, B, h->s_alpha, h->s_start, h->s_goal, h->s_fstate, h->s_istate, -1, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h_alpha, h->s_alpha, (size_t)B * h->T * 3 * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_fstate, h->s_fstate, (size_t)B * FGD_FSTATE * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_istate, h->s_istate, (size_t)B * FGD_ISTATE * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return FGD_OK;
}

int fgd_optimize_host_io(FgdHandle *h, int32_t use_gd, int32_t B, const float *h_alpha_in, float *h_alpha_out,
                         const float *h_start, const float *h_goal, float *h_fstate_out, int32_t *h_istate_out, void *stream)
{
    if (!h || B < 0 || (B > 0 && (!h_alpha_in || !h_alpha_out || !h_start || !h_goal || !h_fstate_out || !h_istate_out)))
        return FGD_ERR_INVALID_ARGUMENT;
    if (B == 0) return FGD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // Zero-copy: when every buffer is page-locked host memory (cudaHostAlloc / cudaHostRegister / torch pin_memory),
    // the persistent kernel reads each trajectory straight from the caller's input buffer when a team picks it up and
    // writes alpha / state straight into the output buffers when the team retires it - PCIe traffic overlaps the
    // arithmetic of the other teams and there is no staging copy, no memset and no separate D2H pass.
    // FGD_HOST_IO=copy forces the staged path (A/B measurements).
    {
        static const bool force_copy = [] { const char *e = std::getenv("FGD_HOST_IO"); return e && std::strcmp(e, "copy") == 0; }();
        void *d_in = nullptr, *d_out = nullptr, *d_s = nullptr, *d_g = nullptr, *d_f = nullptr, *d_i = nullptr;
        const size_t nb_a = (size_t)B * h->T * 3 * 4, nb_sg = (size_t)B * 3 * 4;
        if (!force_copy && mapped_host_range(h_alpha_in, nb_a, &d_in) && mapped_host_range(h_alpha_out, nb_a, &d_out) &&
            mapped_host_range(h_start, nb_sg, &d_s) && mapped_host_range(h_goal, nb_sg, &d_g) &&
            mapped_host_range(h_fstate_out, (size_t)B * FGD_FSTATE * 4, &d_f) && mapped_host_range(h_istate_out, (size_t)B * FGD_ISTATE * 4, &d_i)) {
            int rc = run_optimize(h, use_gd ? 1 : 0, B, (float *)d_out, (const float *)d_s, (const float *)d_g, (float *)d_f, (int *)d_i,
                                  -1, st, (const float *)d_in);
            if (rc) return rc;
            h->zero_copy_calls += 1;
            CK(cudaStreamSynchronize(st));
            return FGD_OK;
        }
    }
    int rc = ensure_scratch(h, B);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->s_alpha, h_alpha_in, (size_t)B * h->T * 3 * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->s_start, h_start, (size_t)B * 3 * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->s_goal, h_goal, (size_t)B * 3 * 4, cudaMemcpyHostToDevice, st));
    // the loop state is write-only for a fresh run (DevParams::fresh): no memset, no upload
    rc = run_optimize(h, use_gd ? 1 : 0, B, h->s_alpha, h->s_start, h->s_goal, h->s_fstate, h->s_istate, -1, st, h->s_alpha);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h_alpha_out, h->s_alpha, (size_t)B * h->T * 3 * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_fstate_out, h->s_fstate, (size_t)B * FGD_FSTATE * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_istate_out, h->s_istate, (size_t)B * FGD_ISTATE * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return FGD_OK;
}

int fgd_argmin_per_problem(FgdHandle *h, int32_t n_problems, int32_t n_restarts, const float *d_fstate, const int32_t *d_istate,
                           int32_t index_offset, int32_t problem_stride, float *d_best_cost, int32_t *d_best_index, int64_t *d_best_key,
                           void *stream)
{
    if (!h || n_problems < 0 || n_restarts < 1 || (n_problems > 0 && (!d_fstate || !d_istate || (!d_best_cost && !d_best_index && !d_best_key))))
        return FGD_ERR_INVALID_ARGUMENT;
    if (n_problems == 0) return FGD_OK;
    if (problem_stride <= 0) problem_stride = n_restarts;
    if ((long long)index_offset + (long long)(n_problems - 1) * problem_stride + n_restarts - 1 > 0x7fffffffLL) return FGD_ERR_INVALID_ARGUMENT;
    const int block = 256, per = block / 32;
    const int grid = (n_problems + per - 1) / per;
    fgd_argmin_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(n_problems, n_restarts, d_fstate, d_istate, index_offset, problem_stride,
                                                                d_best_cost, d_best_index, reinterpret_cast<long long *>(d_best_key));
    CK(cudaGetLastError());
    h->launches += 1;
    return FGD_OK;
}

int fgd_set_init_basis(FgdHandle *h, const float *h_u, const float *h_w, const float *h_jinv)
{
    if (!h || !h_u || !h_w || !h_jinv) return FGD_ERR_INVALID_ARGUMENT;
    const int T = h->T;
    std::vector<float> buf((size_t)2 * T + 9);
    std::memcpy(buf.data(), h_u, (size_t)T * 4);
    std::memcpy(buf.data() + T, h_w, (size_t)T * 4);
    std::memcpy(buf.data() + 2 * T, h_jinv, 9 * 4);
    if (!h->d_init) CK(cudaMalloc(&h->d_init, buf.size() * 4));
    CK(cudaMemcpy(h->d_init, buf.data(), buf.size() * 4, cudaMemcpyHostToDevice));
    return FGD_OK;
}

int fgd_init_trajectory(FgdHandle *h, int32_t B, const float *d_start, const float *d_goal, float *d_alpha, void *stream)
{
    if (!h || B < 0 || (B > 0 && (!d_start || !d_goal || !d_alpha)) || !h->d_init) return FGD_ERR_INVALID_ARGUMENT;
    if (B == 0) return FGD_OK;
    const long long n = (long long)B * h->T;
    long long blocks = (n + 255) / 256, cap = (long long)h->num_sms * 8;
    fgd_init_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(B, h->T, h->d_init, h->d_init + 2 * h->T,
                                                                                         d_start, d_goal, d_alpha);
    CK(cudaGetLastError());
    h->launches += 1;
    return FGD_OK;
}

int fgd_launch_geometry(const FgdHandle *h, int32_t B, int32_t *grid, int32_t *block, int32_t *smem_bytes, int32_t *warps_per_trajectory)
{
    if (!h || B < 1) return FGD_ERR_INVALID_ARGUMENT;
    const Geometry g = geometry(h, B, h->obs_count);
    if (grid) *grid = g.grid;
    if (block) *block = g.block;
    if (smem_bytes) *smem_bytes = g.smem;
    if (warps_per_trajectory) *warps_per_trajectory = h->WPT;
    return FGD_OK;
}

int64_t fgd_kernel_launches(const FgdHandle *h) { return h ? h->launches : 0; }
int64_t fgd_zero_copy_calls(const FgdHandle *h) { return h ? h->zero_copy_calls : 0; }
int64_t fgd_speculative_launches(const FgdHandle *h) { return h ? h->spec_launches : 0; }

#if defined(FGD_DEBUG_MARK) || defined(FGD_PHASE_CLOCKS)
int *fgd_debug_buffer(FgdHandle *h) { return h->h_dbg; }
#endif

int fgd_measure_fp32_peak(FgdHandle *h, double *tflops_out, void *stream)
{
    if (!h || !tflops_out) return FGD_ERR_INVALID_ARGUMENT;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = h->num_sms * 8, iters = 4096;
    return measure_peak(h, st, (double)grid * 256 * (double)iters * 16 * 8 * 2,
                        [&](float seed, float *sink) { fgd_ffma_peak_kernel<<<grid, 256, 0, st>>>(iters, seed, sink); }, tflops_out);
}

int fgd_measure_mufu_peak(FgdHandle *h, double *trcp_out, void *stream)
{
    if (!h || !trcp_out) return FGD_ERR_INVALID_ARGUMENT;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = h->num_sms * 8, iters = 1024;
    return measure_peak(h, st, (double)grid * 256 * (double)iters * 8 * 8,
                        [&](float seed, float *sink) { fgd_mufu_peak_kernel<<<grid, 256, 0, st>>>(iters, seed, sink); }, trcp_out);
}

}  // extern "C"
