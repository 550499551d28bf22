// fgd_device.cuh -- device-side building blocks of the batched FGD iteration (sm_100a).
//
// Mapping.  A trajectory is owned by a TEAM of WPT warps
// (WPT = 1 for T <= 64, 2 for T <= 128, 4 for T <= 256).  Team thread i
// (i = 32 * warp_in_team + lane) owns the R = 2 adjacent time samples t = 2i, 2i+1,
// handled as ONE packed FP32 row pair (FFMA2 / FADD2 / FMUL2).  The RKHS contraction
// produces rows t of q = K alpha J and v = dK alpha J in the thread that then does
// forward kinematics, the obstacle potential and the penalty terms for those samples --
// no shared-memory round trip between the two.  K and dK are re-laid out per column k so
// that a thread fetches its entries of column k with one access: from TENSOR MEMORY
// (tcgen05.ld, the default: one table copy per SM shared by all its teams), from shared
// memory or from L2 (the A/B variants and the evaluation kernel).  The operand rows
// (alpha' or the q/v-gradients) sit in a per-team shared buffer and enter the FFMA2 as
// broadcast scalars.
// Reductions over t: lane-serial over the 2 rows, a 5-step xor butterfly inside each
// warp (three sums share one transposed butterfly, the maximum is one integer REDUX),
// then -- WPT > 1 -- the warp partials are exchanged through shared memory and combined
// as (p0 + p1) + (p2 + p3): the order the mirror oracle reproduces.  All trajectory-level
// decisions are computed redundantly by every thread of the team from the same reduced
// values, so the team stays convergent (team barrier = __syncwarp for WPT = 1, a named
// barrier per team for WPT > 1, the CTA barrier when a team owns its CTA).
// Lanes of a single-warp team that own no sample (T = 50: 7 of 32) take a share of the
// obstacle loop in many-obstacle scenes (share_split, cost_phase).
//
// Compiled with -fmad=false: every fused multiply-add is an explicit fmaf() / fma2(), so
// the operation sequence is the documented one (bit-exact against the oracle in
// strict-math mode).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fgd {

constexpr unsigned FULL = 0xffffffffu;
constexpr int R = 2;             // adjacent time samples per thread (one packed row pair)

enum Kind : int { K_IDLE = 0, K_EVAL0 = 1, K_CAND = 2, K_BACK = 3 };

struct DevParams {
    int T, TP, n_obs, max_inner, max_outer, max_bls, cvdl, mode, budget, B;
    float lam_sg0, lam_jl0, lam_inc, lam_max, lam_reg, eps_loop, eps_pos, eps_vel;
    float thr_pos, thr_vel;  // smallest floats whose IEEE square root reaches eps_pos / eps_vel:  sqrt(x) < eps  <=>  x < thr  (host, fgd_create)
    float bls_lr0, bls_alpha, bls_bp, bls_bm;
    float qmax, qmin, vmax;
    float link[3];
    float J[9];
    float gd_lr[16];
    // derived on the host, rounded once to FP32 (same expressions as the oracle)
    float oml, inv_T, w_avg, mean_q, inv_std, inv_std2, inv_vmax, inv_vmax2, q_hi, q_lo, v_hi, fT;
    const float *KD;         // [T][TP/2][4]: K[t][k], K[t+1][k], dK[t][k], dK[t+1][k] of team thread t/2 (zero padded rows >= T)
    const float *KO;         // [T][TP/2][2]: K entries only (dense half of the backward contraction)
    const float *DO;         // [T][TP/2][2]: dK entries only (T > 128: K lives in tensor memory, dK streams from L2)
    const float *obs;        // [n_obs][2]
    float *alpha;            // [B][T][3]  final alpha rows are written here
    const float *alpha_in;   // [B][T][3]  initial alpha rows are read from here (== alpha for the in-place calls)
    int fresh;               // 1: every trajectory starts from the zero loop state (istate / fstate are write-only)
    const float *start, *goal;
    float *fstate;
    int *istate;
    unsigned *queue;
    int *dbg;                // optional host-mapped progress markers (debug builds only)
    // live obstacle updates (fgd_optimize_live, poll_every > 0): the ring of published obstacle sets
    const int *obs_meta;     // [0]: latest published generation; [2 + 2s], [3 + 2s]: generation and count held by ring slot s
    const float *obs_ring;   // [FGD_OBS_RING][obs_cap][2]
    int obs_cap, poll_every;
    int *switch_log;         // optional [B][FGD_SWITCH_LOG][2]
};

struct EvalPtrs {
    float lam_sg, lam_jl;
    float *loss, *toc, *grad, *q, *v;
    int *fulfilled;
};

// team-uniform per-trajectory scalars (one copy per thread, identical across the team)
struct Slot {
    int traj, status, outer, inner, inner_total, cand_evals, accepts, ful, j, done_iters;
    unsigned hash;
    float lam_sg, lam_jl, lr, loss, toc, alpha_norm, last_new;
    float start[3], goal[3];
};

// Exchange scratch of one team in shared memory (WPT > 1).  Every exchange site has its own
// words; a site is written at most once per loop trip and every trip passes at least two team
// barriers (around the contraction), so a site is never overwritten while a team mate still reads it.
constexpr int XCH_COST = 0;      // [WPT][8]: max, argmax, sum cost, sum jp, sum jv, limits ok
constexpr int XCH_ENDS = 32;     // ssp0, ssv0, sspT, ssvT
constexpr int XCH_NORM = 36;     // [WPT]
constexpr int XCH_ANORM = 40;    // [WPT]
constexpr int XCH_NZ = 44;       // [WPT][R]
constexpr int XCH_FETCH = 52;    // queue index
constexpr int XCH_WORDS = 56;

// geometry of one trajectory team
template <int WPT>
struct Team {
    int lane, wit, tl;       // lane in warp, warp in team, thread in team
    int bar;                 // WPT > 1: named barrier of this team (several teams per CTA), 0 = the CTA barrier (one team per CTA)
    float *xch;
    __device__ __forceinline__ Team(float *xch_, int bar_ = 0)
    {
        lane = threadIdx.x & 31;
        wit = (WPT == 1) ? 0 : ((threadIdx.x >> 5) & (WPT - 1));
        tl = wit * 32 + lane;
        bar = bar_;
        xch = xch_;
    }
    __device__ __forceinline__ void sync() const
    {
        if constexpr (WPT == 1) __syncwarp();
        else if (bar == 0) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(WPT * 32) : "memory");
    }
};

// warp collectives (always executed by all 32 lanes)
__device__ __forceinline__ float wsum(float v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = v + __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float wmax(float v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
// Maximum over the warp of NON-NEGATIVE, non-NaN floats: their bit patterns order like unsigned integers, so one
// integer warp reduction (REDUX.MAX.U32) returns the bits of the maximum - the same value as the fmaxf butterfly.
__device__ __forceinline__ float wmax_nonneg(float v)
{
    return __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(v)));
}
// Three warp sums at once, each bit-identical to wsum(): the xor butterfly adds the same pairs in every lane
// (u_l = v_l + v_{l^16}, w_l = u_l + u_{l^8}, ...), so a lane only has to carry the partial of ONE value once the
// halves of the warp split the work: after the 16-step lanes 0-15 carry a and b, lanes 16-31 carry c; after the 8-step
// lanes 0-7 carry a, 8-15 b, 16-31 c; the 4-, 2- and 1-steps are plain butterflies and three index shuffles hand the
// totals to every lane.  9 shuffles and 6 additions instead of 15 and 15.
__device__ __forceinline__ void wsum3(float &a, float &b, float &c)
{
    const int lane = threadIdx.x & 31;
    const bool lo = lane < 16, b3 = (lane & 8) != 0;
    // 16-step
    const float r1 = __shfl_xor_sync(FULL, lo ? c : a, 16);      // lo receives the partner's a, hi the partner's c
    const float r2 = __shfl_xor_sync(FULL, b, 16);               // lo receives the partner's b
    const float x1 = (lo ? a : c) + r1;
    const float y1 = b + r2;                                     // meaningful in lanes 0-15
    // 8-step: lanes 0-7 keep a and send b, lanes 8-15 keep b and send a, lanes 16-31 keep and send c
    const bool keep_y = lo && b3;
    const float keep = keep_y ? y1 : x1;
    const float send = lo ? (b3 ? x1 : y1) : x1;
    float z = keep + __shfl_xor_sync(FULL, send, 8);
    z = z + __shfl_xor_sync(FULL, z, 4);
    z = z + __shfl_xor_sync(FULL, z, 2);
    z = z + __shfl_xor_sync(FULL, z, 1);
    a = __shfl_sync(FULL, z, 0);
    b = __shfl_sync(FULL, z, 8);
    c = __shfl_sync(FULL, z, 16);
}
// cross-warp combination of warp partials p[0..WPT-1] in butterfly order
template <int WPT>
__device__ __forceinline__ float combine_sum(const float *p)
{
    if constexpr (WPT == 1) return p[0];
    else if constexpr (WPT == 2) return p[0] + p[1];
    else return (p[0] + p[1]) + (p[2] + p[3]);
}

// sum over the whole team of a per-thread partial (one exchange, one team barrier for WPT > 1)
template <int WPT>
__device__ __forceinline__ float tsum(const Team<WPT> &G, float v, int site)
{
    v = wsum(v);
    if constexpr (WPT == 1) return v;
    else {
        if (G.lane == 0) G.xch[site + G.wit] = v;
        G.sync();
        float p[WPT];
#pragma unroll
        for (int w = 0; w < WPT; ++w) p[w] = G.xch[site + w];
        return combine_sum<WPT>(p);
    }
}

__device__ __forceinline__ float ss3(float a, float b, float c) { return fmaf(c, c, fmaf(b, b, a * a)); }

// ---------------------------------------------------------------------------
// Packed FP32 (sm_100 FFMA2 / FMUL2 / FADD2).  A lane's R adjacent time samples
// are handled as R/2 row pairs: .x = row 2p, .y = row 2p+1.  Every packed
// instruction is two independent IEEE-754 round-to-nearest operations, so the
// results are bit-identical to the scalar sequence of the mirror oracle while
// the instruction count per row halves.  Caveat (seen in SASS): ptxas contracts an
// __fmul2_rn feeding an __fadd2_rn into one FFMA2 even under -fmad=false, so a
// product that is added to something is always written as an explicit fma2() here
// and as fmaf() in the oracle; mul2 results only feed multiplies, selects or stores.  Scalars (constants, obstacle
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

// Reciprocals of U packed pairs m = 1 + |f - o|^2 (>= 1, or NaN).  Fast: MUFU.RCP.  Strict: the IEEE reciprocal.
// __frcp_rn is MUFU.RCP + one Newton step (e = fma(r, m, -1), r' = fma(r, -e, r)) behind an exponent check that sends
// m >= 2^126, infinities, NaNs, zeros and denormals to a slow path - per value: ~9 scalar instructions and a branch, which
// made the strict obstacle loop 2.4x slower than the fast one.  Here m >= 1, so the only slow-path case is m >= 2^126:
// ONE check per block (the maximum of the block's values) guards the same Newton step in packed form (two FFMA2 per pair,
// the same operations as the fast path of __frcp_rn: bit-identical), a block with an absurdly distant obstacle takes
// __frcp_rn.  Both branches return the correctly rounded reciprocal.
template <int U, bool STRICT>
__device__ __forceinline__ void rcp_block(f2 (&m)[U])
{
    if constexpr (!STRICT) {
#pragma unroll
        for (int u = 0; u < U; ++u) m[u] = mk2(rcp<false>(m[u].x), rcp<false>(m[u].y));
    } else {
        float mx = fmaxf(m[0].x, m[0].y);
#pragma unroll
        for (int u = 1; u < U; ++u) mx = fmaxf(mx, fmaxf(m[u].x, m[u].y));
        if (mx < 8.507059173e+37f) {                    // 2^126
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const f2 r = mk2(rcp<false>(m[u].x), rcp<false>(m[u].y));
                const f2 e = fma2(r, m[u], bc2(-1.0f));
                m[u] = fma2(r, neg2(e), r);
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) m[u] = mk2(__frcp_rn(m[u].x), __frcp_rn(m[u].y));
        }
    }
}

// x / T for the means over the time samples.  Strict: the IEEE division.  Fast: q0 = x * (1/T), one exact-remainder
// correction (Markstein) - equal to the IEEE quotient for finite x (tests/test_oracle_mirror.py checks the sequence
// against true division), 3 instructions instead of ~16 on the scalar critical path.
template <bool STRICT>
__device__ __forceinline__ float div_T(const DevParams &p, float x)
{
    if constexpr (STRICT) {
        return x / p.fT;
    } else {
        const float q0 = x * p.inv_T;
        return fmaf(fmaf(-q0, p.fT, x), p.inv_T, q0);
    }
}

// sin and cos of two angles: Cody-Waite reduction by pi/2 (3 constants) and the
// cephes minimax polynomials on [-pi/4, pi/4]; <= 2 ulp for the angles a 3-joint
// arm with limits [-1, 2] rad produces.  Same operation sequence as the oracle.
// The quadrant index comes from the magic-number rounding t = fma(x, 2/pi, 1.5 * 2^23): the low mantissa bits of t
// are round-to-nearest-even(x * 2/pi) in two's complement and j = t - 1.5 * 2^23 is that integer as a float - two
// packed operations instead of an FRND and an F2I per angle and row (both quarter-rate).  Quadrant n = bits & 3:
// swap sin / cos when n is odd, negate sin for n >= 2 (bit 1 of n), negate cos for n = 1, 2 (bit 1 of n + 1);
// the negations are sign-bit flips done with integer logic.
constexpr float SINCOS_MAGIC = 12582912.0f;       // 1.5 * 2^23
__device__ __forceinline__ void quadrant(unsigned n, float sn, float cs, float &S, float &C)
{
    const bool sw = (n & 1u) != 0u;
    const float s0 = sw ? cs : sn, c0 = sw ? sn : cs;
    S = __uint_as_float(__float_as_uint(s0) ^ ((n << 30) & 0x80000000u));
    C = __uint_as_float(__float_as_uint(c0) ^ (((n + 1u) << 30) & 0x80000000u));
}

__device__ __forceinline__ void sincos_cw2(f2 x, f2 &S, f2 &C)
{
    const f2 t = fma2(x, bc2(6.366197467e-01f), bc2(SINCOS_MAGIC));
    const f2 j = add2(t, bc2(-SINCOS_MAGIC));
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
    quadrant(__float_as_uint(t.x), sn.x, cs.x, S.x, C.x);
    quadrant(__float_as_uint(t.y), sn.y, cs.y, S.y, C.y);
}

// where a kernel instance keeps the K / dK operand tables
enum KSource : int { K_L2 = 0, K_SMEM = 1, K_TMEM = 2 };

// ---------------------------------------------------------------------------
// Tensor memory (TMEM, 128 lanes x 512 columns x 32 bit per SM) as a per-lane table store (T <= 64,
// single-warp teams).  Thread `lane` of every warp needs, per column k of the contraction, its own four
// table entries KD[k][lane] = { K[2l][k], K[2l+1][k], dK[2l][k], dK[2l+1][k] }.  From shared memory that is a
// 512 B LDS.128 per warp and k - measured ~6 data-pipe cycles - and the shared-memory pipe was the busiest unit
// of the kernel (75 % of peak, ncu).  TMEM lane l / column 4k..4k+3 holds the same four words: one
// tcgen05.ld.32x32b (SASS LDTM) hands every thread its entries through the tensor-memory datapath, which is not
// the shared-memory pipe, sustains > 700 B/cycle/SM (profiles/scripts/probe_tmem.cu) and has a 12-cycle latency.  The
// table is written once per CTA with tcgen05.st; warp w reads the lane quadrant 32 (w % 4), so warps w and
// w + 4 share one copy.  No tensor-core instruction is involved: the arithmetic stays FP32 FFMA2.
// The registers of a tcgen05.ld are only valid after tcgen05.wait::ld; the wait takes them as in/out operands
// so that the compiler cannot move a consumer above it.
// ---------------------------------------------------------------------------
constexpr int TMEM_COLS = 256;        // 4 columns per k, T <= 64; two resident CTAs per SM own all 512 columns

__device__ __forceinline__ void tmem_ld16(unsigned ta, unsigned (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(ta));
}
__device__ __forceinline__ void tmem_wait16(unsigned (&r)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]) :: "memory");
}
__device__ __forceinline__ void tmem_ld4(unsigned ta, unsigned (&r)[4])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(ta));
}
__device__ __forceinline__ void tmem_wait4(unsigned (&r)[4])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]) :: "memory");
}
__device__ __forceinline__ void tmem_ld2(unsigned ta, unsigned (&r)[2])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(ta));
}
__device__ __forceinline__ void tmem_wait2(unsigned (&r)[2])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]) :: "memory");
}
__device__ __forceinline__ void tmem_ld8(unsigned ta, unsigned (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(ta));
}
__device__ __forceinline__ void tmem_wait8(unsigned (&r)[8])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]) :: "memory");
}
__device__ __forceinline__ f2 u2f2(unsigned a, unsigned b) { return make_float2(__uint_as_float(a), __uint_as_float(b)); }

// Forward contraction with the operand table in TMEM: same terms, same order (k ascending, one fma per term) as
// contract<>.  tk = TMEM address of this warp's lane quadrant, column 0.
// TC > 0: T is the compile-time constant TC and the loops are fully unrolled (immediate TMEM columns and operand offsets,
// no loop bookkeeping); TC = 0: runtime T.
template <bool SAME, int TC>
__device__ __forceinline__ void contract_tm(unsigned tk, int T, const float4 *__restrict__ x1, const float4 *__restrict__ x2,
                                            f2 (&y1)[3], f2 (&y2)[3])
{
#pragma unroll
    for (int a = 0; a < 3; ++a) { y1[a] = bc2(0.0f); y2[a] = bc2(0.0f); }
    int k = 0;
#pragma unroll (TC > 0 ? 64 : 1)
    for (; k + 4 <= T; k += 4) {
        unsigned r[16];
        tmem_ld16(tk + 4 * k, r);
        float4 xa[4], xb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { xa[u] = x1[k + u]; xb[u] = xa[u]; if constexpr (!SAME) xb[u] = x2[k + u]; }
        tmem_wait16(r);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const f2 kk = u2f2(r[4 * u], r[4 * u + 1]), dk = u2f2(r[4 * u + 2], r[4 * u + 3]);
            y1[0] = fma2(kk, bc2(xa[u].x), y1[0]);
            y1[1] = fma2(kk, bc2(xa[u].y), y1[1]);
            y1[2] = fma2(kk, bc2(xa[u].z), y1[2]);
            y2[0] = fma2(dk, bc2(xb[u].x), y2[0]);
            y2[1] = fma2(dk, bc2(xb[u].y), y2[1]);
            y2[2] = fma2(dk, bc2(xb[u].z), y2[2]);
        }
    }
#pragma unroll (TC > 0 ? 4 : 1)
    for (; k < T; ++k) {
        unsigned r[4];
        tmem_ld4(tk + 4 * k, r);
        const float4 xa = x1[k];
        float4 xb = xa;
        if constexpr (!SAME) xb = x2[k];
        tmem_wait4(r);
        const f2 kk = u2f2(r[0], r[1]), dk = u2f2(r[2], r[3]);
        y1[0] = fma2(kk, bc2(xa.x), y1[0]);
        y1[1] = fma2(kk, bc2(xa.y), y1[1]);
        y1[2] = fma2(kk, bc2(xa.z), y1[2]);
        y2[0] = fma2(dk, bc2(xb.x), y2[0]);
        y2[1] = fma2(dk, bc2(xb.y), y2[1]);
        y2[2] = fma2(dk, bc2(xb.z), y2[2]);
    }
}

// Backward contraction with the operand table in TMEM (see contract_back<>): K G_q dense, dK (-G_v) over the
// flagged rows only, both in ascending k.  Single-warp teams.
template <int TC>
__device__ __forceinline__ void contract_back_tm(unsigned tk, int T, const float4 *__restrict__ xa_rows, const float4 *__restrict__ xb_rows,
                                                 const unsigned (&nz)[1][R], f2 (&y1)[3], f2 (&y2)[3])
{
#pragma unroll
    for (int a = 0; a < 3; ++a) { y1[a] = bc2(0.0f); y2[a] = bc2(0.0f); }
    int k = 0;
#pragma unroll (TC > 0 ? 64 : 1)
    for (; k + 4 <= T; k += 4) {
        unsigned r[16];
        tmem_ld16(tk + 4 * k, r);
        float4 xa[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) xa[u] = xa_rows[k + u];
        tmem_wait16(r);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const f2 kk = u2f2(r[4 * u], r[4 * u + 1]);
            y1[0] = fma2(kk, bc2(xa[u].x), y1[0]);
            y1[1] = fma2(kk, bc2(xa[u].y), y1[1]);
            y1[2] = fma2(kk, bc2(xa[u].z), y1[2]);
        }
    }
#pragma unroll (TC > 0 ? 4 : 1)
    for (; k < T; ++k) {
        unsigned r[2];
        tmem_ld2(tk + 4 * k, r);
        const float4 xa = xa_rows[k];
        tmem_wait2(r);
        const f2 kk = u2f2(r[0], r[1]);
        y1[0] = fma2(kk, bc2(xa.x), y1[0]);
        y1[1] = fma2(kk, bc2(xa.y), y1[1]);
        y1[2] = fma2(kk, bc2(xa.z), y1[2]);
    }
    // dK (-G_v): rows 0 and T-1 carry the start/goal velocity terms in every iteration (trajectory.py:207-212), so they are
    // visited unconditionally as straight-line code (a row that happens to be all zero adds exact zeros: fma(dv, 0, y) == y,
    // y never being -0); the rows in between only where the velocity limit is violated.  k ascending throughout.
    auto dk_row = [&](const int kz) {
        unsigned r[2];
        tmem_ld2(tk + 4 * kz + 2, r);
        const float4 xb = xb_rows[kz];
        tmem_wait2(r);
        const f2 dv = u2f2(r[0], r[1]);
        y2[0] = fma2(dv, bc2(xb.x), y2[0]);
        y2[1] = fma2(dv, bc2(xb.y), y2[1]);
        y2[2] = fma2(dv, bc2(xb.z), y2[2]);
    };
    const int lT = (T - 1) / R, rT = (T - 1) % R;
    unsigned m0 = nz[0][0] & ~1u, m1 = nz[0][1];           // row 0 = lane 0, slot 0
    if (rT == 0) m0 &= ~(1u << lT); else m1 &= ~(1u << lT);
    dk_row(0);
    unsigned any = m0 | m1;
    while (any) {                                   // warp-uniform: ascending lane, then ascending slot = ascending k
        const int l = __ffs(any) - 1;
        any &= any - 1;
        if ((m0 >> l) & 1u) dk_row(l * R);
        if ((m1 >> l) & 1u) dk_row(l * R + 1);
    }
    dk_row(T - 1);
}

// ---------------------------------------------------------------------------
// Multi-warp teams (T > 64) with the tables in tensor memory: ONE 16-warp CTA per SM holds all 512 TMEM columns and its
// 16 / WPT teams share them - warp w reads lane quadrant w % 4, which holds the entries of team thread
// 32 (w % WPT) + lane for every team alike.  Reading the tables through the L1 / shared-memory pipe costs 4 wavefronts
// per warp and column k (512 B) against 12 FMA-pipe cycles of arithmetic, i.e. 80 data-pipe cycles against 48 FMA cycles
// per column over the SM's 16 warps: the contraction ran at ~60 % of the obstacle loop's efficiency (phase clocks,
// profiles/r02f_phase_clocks_loaded.txt).  The TMEM read path is a different one.
//   WPT = 2 (T <= 128): K and dK (4 columns per k = 4 T <= 512 columns) - contract_tm / contract_back_mw<4>.
//   WPT = 4 (T <= 256): K alone (2 columns per k); dK streams from the compact table DO in L2, 2 wavefronts per warp
//   and k, and only for the forward contraction and the flagged rows of the backward one.
// Same terms in the same order as contract<> / contract_back<>: bit-identical results.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void contract_tm_k(unsigned tk, const float *__restrict__ dd, const int stride, int T, const float4 *__restrict__ x,
                                              f2 (&y1)[3], f2 (&y2)[3])
{
#pragma unroll
    for (int a = 0; a < 3; ++a) { y1[a] = bc2(0.0f); y2[a] = bc2(0.0f); }
    int k = 0;
#pragma unroll 2
    for (; k + 4 <= T; k += 4) {
        unsigned r[8];
        tmem_ld8(tk + 2 * k, r);
        float2 dv[4];
        float4 xa[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { dv[u] = __ldg(reinterpret_cast<const float2 *>(dd + (size_t)(k + u) * stride)); xa[u] = x[k + u]; }
        tmem_wait8(r);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const f2 kk = u2f2(r[2 * u], r[2 * u + 1]), dk = mk2(dv[u].x, dv[u].y);
            y1[0] = fma2(kk, bc2(xa[u].x), y1[0]);
            y1[1] = fma2(kk, bc2(xa[u].y), y1[1]);
            y1[2] = fma2(kk, bc2(xa[u].z), y1[2]);
            y2[0] = fma2(dk, bc2(xa[u].x), y2[0]);
            y2[1] = fma2(dk, bc2(xa[u].y), y2[1]);
            y2[2] = fma2(dk, bc2(xa[u].z), y2[2]);
        }
    }
    for (; k < T; ++k) {
        unsigned r[2];
        tmem_ld2(tk + 2 * k, r);
        const float2 dv = __ldg(reinterpret_cast<const float2 *>(dd + (size_t)k * stride));
        const float4 xa = x[k];
        tmem_wait2(r);
        const f2 kk = u2f2(r[0], r[1]), dk = mk2(dv.x, dv.y);
        y1[0] = fma2(kk, bc2(xa.x), y1[0]);
        y1[1] = fma2(kk, bc2(xa.y), y1[1]);
        y1[2] = fma2(kk, bc2(xa.z), y1[2]);
        y2[0] = fma2(dk, bc2(xa.x), y2[0]);
        y2[1] = fma2(dk, bc2(xa.y), y2[1]);
        y2[2] = fma2(dk, bc2(xa.z), y2[2]);
    }
}

// Backward contraction, multi-warp teams, K from tensor memory (CPK columns per k: 4 = K and dK interleaved, 2 = K alone);
// the flagged dK rows from tensor memory (CPK = 4) or from DO (CPK = 2).
template <int WPT, int CPK>
__device__ __forceinline__ void contract_back_mw(unsigned tk, const float *__restrict__ dd, int T, const float4 *__restrict__ xa_rows,
                                                 const float4 *__restrict__ xb_rows, const unsigned (&nz)[WPT][R], f2 (&y1)[3], f2 (&y2)[3])
{
    constexpr int SO = WPT * 32 * R;
#pragma unroll
    for (int a = 0; a < 3; ++a) { y1[a] = bc2(0.0f); y2[a] = bc2(0.0f); }
    int k = 0;
    if constexpr (CPK == 2) {
#pragma unroll 2
        for (; k + 8 <= T; k += 8) {
            unsigned r[16];
            tmem_ld16(tk + 2 * k, r);
            float4 xa[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) xa[u] = xa_rows[k + u];
            tmem_wait16(r);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const f2 kk = u2f2(r[2 * u], r[2 * u + 1]);
                y1[0] = fma2(kk, bc2(xa[u].x), y1[0]);
                y1[1] = fma2(kk, bc2(xa[u].y), y1[1]);
                y1[2] = fma2(kk, bc2(xa[u].z), y1[2]);
            }
        }
    } else {
#pragma unroll 2
        for (; k + 4 <= T; k += 4) {
            unsigned r[16];
            tmem_ld16(tk + 4 * k, r);
            float4 xa[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) xa[u] = xa_rows[k + u];
            tmem_wait16(r);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const f2 kk = u2f2(r[4 * u], r[4 * u + 1]);
                y1[0] = fma2(kk, bc2(xa[u].x), y1[0]);
                y1[1] = fma2(kk, bc2(xa[u].y), y1[1]);
                y1[2] = fma2(kk, bc2(xa[u].z), y1[2]);
            }
        }
    }
    for (; k < T; ++k) {
        unsigned r[2];
        tmem_ld2(tk + CPK * k, r);
        const float4 xa = xa_rows[k];
        tmem_wait2(r);
        const f2 kk = u2f2(r[0], r[1]);
        y1[0] = fma2(kk, bc2(xa.x), y1[0]);
        y1[1] = fma2(kk, bc2(xa.y), y1[1]);
        y1[2] = fma2(kk, bc2(xa.z), y1[2]);
    }
#pragma unroll
    for (int w = 0; w < WPT; ++w) {
        unsigned any = nz[w][0] | nz[w][1];
        while (any) {                               // team-uniform: ascending thread, then ascending r = ascending k
            const int l = __ffs(any) - 1;
            any &= any - 1;
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                if ((nz[w][rr] >> l) & 1u) {
                    const int kz = (w * 32 + l) * R + rr;
                    const float4 xb = xb_rows[kz];
                    f2 dv;
                    if constexpr (CPK == 4) {
                        unsigned r[2];
                        tmem_ld2(tk + 4 * kz + 2, r);
                        tmem_wait2(r);
                        dv = u2f2(r[0], r[1]);
                    } else {
                        const float2 t = __ldg(reinterpret_cast<const float2 *>(dd + (size_t)kz * SO));
                        dv = mk2(t.x, t.y);
                    }
                    y2[0] = fma2(dv, bc2(xb.x), y2[0]);
                    y2[1] = fma2(dv, bc2(xb.y), y2[1]);
                    y2[2] = fma2(dv, bc2(xb.z), y2[2]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// RKHS contraction for the trajectory of this thread's team:
//   y1[a] = sum_k K [t][k] * x1[k][a]      (trajectory.py:65 / :295)   t = 2*tl, 2*tl+1 packed
//   y2[a] = sum_k dK[t][k] * x2[k][a]
// k ascending, one fma per term.  KD is the interleaved operand table
//   KD[k][tl] = { K[2tl][k], K[2tl+1][k], dK[2tl][k], dK[2tl+1][k] }
// (column stride 2*TP floats), so a thread fetches its entries of column k with one
// 128-bit load and the adjacent rows land in aligned register pairs: one FFMA2 per
// joint, the operand x[k][a] entering as a broadcast scalar.  The table comes from
// shared memory (KS) or, for T > 64, from L2/L1 through the read-only path.
// SAME = true: x1 == x2 (forward evaluation), one operand load per k.
// ---------------------------------------------------------------------------
template <int WPT, int KS, bool SAME>
__device__ __forceinline__ void contract(const float *__restrict__ kd, int T,
                                         const float4 *__restrict__ x1, const float4 *__restrict__ x2,
                                         f2 (&y1)[3], f2 (&y2)[3])
{
    constexpr int STRIDE = 2 * WPT * 32 * R;       // floats per column k
    static_assert(KS == K_L2 || KS == K_SMEM, "the TMEM form is contract_tm");
    constexpr int UNROLL = KS == K_SMEM ? 5 : 8;
#pragma unroll
    for (int a = 0; a < 3; ++a) { y1[a] = bc2(0.0f); y2[a] = bc2(0.0f); }

#pragma unroll UNROLL
    for (int k = 0; k < T; ++k) {
        float4 kv;
        if constexpr (KS == K_SMEM) kv = *reinterpret_cast<const float4 *>(kd + (size_t)k * STRIDE);
        else kv = __ldg(reinterpret_cast<const float4 *>(kd + (size_t)k * STRIDE));
        const float4 xa = x1[k];
        float4 xb = xa;
        if constexpr (!SAME) xb = x2[k];
        const f2 kk = mk2(kv.x, kv.y), dk = mk2(kv.z, kv.w);
        y1[0] = fma2(kk, bc2(xa.x), y1[0]);
        y1[1] = fma2(kk, bc2(xa.y), y1[1]);
        y1[2] = fma2(kk, bc2(xa.z), y1[2]);
        y2[0] = fma2(dk, bc2(xb.x), y2[0]);
        y2[1] = fma2(dk, bc2(xb.y), y2[1]);
        y2[2] = fma2(dk, bc2(xb.z), y2[2]);
    }
}

// ---------------------------------------------------------------------------
// Backward contraction  y1 = K G_q (dense),  y2 = dK (-G_v) (sparse).
// G_v = lam_sg*sgv_g + lam_jl*jv_g is zero except in rows 0 and T-1 and where the
// velocity limit is violated (trajectory.py:207-212, 258-268), so the dK half only
// visits the rows flagged in nz[w][r] (bit l <=> row 2*(32w + l) + r is non-zero).
// Skipped terms are exact zeros, so the result equals the dense sum bit for bit;
// the visited terms are still accumulated in ascending k.
// ---------------------------------------------------------------------------
template <int WPT, int KS>
__device__ __forceinline__ void contract_back(const float *__restrict__ ko, const float *__restrict__ kd, int T,
                                              const float4 *__restrict__ xa_rows, const float4 *__restrict__ xb_rows,
                                              const unsigned (&nz)[WPT][R], f2 (&y1)[3], f2 (&y2)[3])
{
    static_assert(KS == K_L2 || KS == K_SMEM, "the TMEM form is contract_back_tm");
    constexpr int SO = WPT * 32 * R, SD = 2 * WPT * 32 * R;
    constexpr int UNROLL = KS == K_SMEM ? 5 : 8;
#pragma unroll
    for (int a = 0; a < 3; ++a) { y1[a] = bc2(0.0f); y2[a] = bc2(0.0f); }
#pragma unroll UNROLL
    for (int k = 0; k < T; ++k) {
        float2 v;
        if constexpr (KS == K_SMEM) v = *reinterpret_cast<const float2 *>(ko + (size_t)k * SO);
        else v = __ldg(reinterpret_cast<const float2 *>(ko + (size_t)k * SO));
        const float4 xa = xa_rows[k];
        const f2 kk = mk2(v.x, v.y);
        y1[0] = fma2(kk, bc2(xa.x), y1[0]);
        y1[1] = fma2(kk, bc2(xa.y), y1[1]);
        y1[2] = fma2(kk, bc2(xa.z), y1[2]);
    }
#pragma unroll
    for (int w = 0; w < WPT; ++w) {
        unsigned any = nz[w][0] | nz[w][1];
        while (any) {                               // team-uniform: ascending thread, then ascending r = ascending k
            const int l = __ffs(any) - 1;
            any &= any - 1;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if ((nz[w][r] >> l) & 1u) {
                    const int k = (w * 32 + l) * R + r;
                    const float4 xb = xb_rows[k];
                    const float2 *col = reinterpret_cast<const float2 *>(kd + (size_t)k * SD + R);
                    const f2 dv = KS == K_SMEM ? col[0] : __ldg(col);
                    y2[0] = fma2(dv, bc2(xb.x), y2[0]);
                    y2[1] = fma2(dv, bc2(xb.y), y2[1]);
                    y2[2] = fma2(dv, bc2(xb.z), y2[2]);
                }
            }
        }
    }
}

// Per-thread row pair kept between the cost phase and the gradient phase.  ARM (whole-arm obstacle cost,
// DevBlog-Theme/blog-post.html:505-513): the potential's gradient at the three joint positions fk_1, fk_2,
// fk_3 (robot.py:39-72); otherwise only at the end effector fk_3 = fk (index 0).
template <bool ARM>
struct Rows {
    f2 q[3], v[3], sn[3], cs[3];
    f2 gx[ARM ? 3 : 1], gy[ARM ? 3 : 1];
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

// U obstacles against the row pair (environment.py:32-58) in the scaled form m = 2 den = 1 + |f - o|^2:
// sr += 1/m, sx += dx/m^2, sy += dy/m^2; the powers of two are exact and folded into the per-sample
// constants (0.8 * 2, -0.8 * 4), which saves one FP32 operation per (sample, obstacle) pair.  Three stages (distances -> reciprocals -> accumulation) so that U
// independent packed chains are in flight across the MUFU latency; the accumulation stage visits the
// obstacles in ascending order (the oracle's summation order).
template <int U, bool STRICT>
__device__ __forceinline__ void obstacle_block(const float2 *__restrict__ obs, const f2 x, const f2 y, f2 &sr, f2 &sx, f2 &sy)
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
    f2 dx[U], dy[U], rr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        dx[u] = add2(x, bc2(-ob[u].x));
        dy[u] = add2(y, bc2(-ob[u].y));
        rr[u] = fma2(dy[u], dy[u], fma2(dx[u], dx[u], bc2(1.0f)));
    }
    rcp_block<U, STRICT>(rr);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        sr = add2(sr, rr[u]);
        const f2 r2 = mul2(rr[u], rr[u]);
        sx = fma2(r2, dx[u], sx);
        sy = fma2(r2, dy[u], sy);
    }
}

// Software-pipelined form of the obstacle loop (end-effector cost): a PAIR of obstacles per stage.  obstacle_pair_head
// computes the distances and issues the reciprocals of a pair, obstacle_pair_tail accumulates a pair; the loop runs the head
// of pair p+1 before the tail of pair p, so the MUFU latency of a pair is covered by the independent arithmetic of the next
// one even for a lone warp (in the block form the accumulation stage waits for the reciprocals issued just before it:
// `wait` was the top stall of the obstacle-bound shapes, FMA pipe 70 %).  Same operations, same ascending accumulation
// order as obstacle_block<>: bit-identical.  Used (cost_phase<..., PIPE>) by the multi-warp-team kernels (T > 64): config 3
// +9.4 % (profiles/r02j_*).  The single-warp-team kernels keep the block form: the pipelined loop measured 0.7 % (c2) to
// 3.7 % (c4, 256 obstacles) slower in the T = 50 instances (at their register pressure ptxas moves the tail of a pair
// right behind its own reciprocals again, seen in SASS) and 5.5 % slower in the runtime-T LIVE instance, where the
// schedule survives (profiles/r02k_*).
#ifndef FGD_OBS_PIPE
#define FGD_OBS_PIPE 1
#endif
#ifndef FGD_UNROLL_PIPE
#define FGD_UNROLL_PIPE 2        // trips of the pipelined loop per loop iteration (multi-warp teams): 2 measured -3.6 % on config 3, 4 only -2.4 % (profiles/r02zm_unroll.txt)
#endif
#ifndef FGD_UNROLL_R4
#define FGD_UNROLL_R4 1          // blocks of four per loop iteration in the shared obstacle loop (single-warp teams, many obstacles)
#endif
constexpr int UNROLL_PIPE = FGD_UNROLL_PIPE, UNROLL_R4 = FGD_UNROLL_R4;
struct ObsPair { f2 dx[2], dy[2], rr[2]; };

template <bool STRICT>
__device__ __forceinline__ void obstacle_pair_head(const float2 *__restrict__ obs, const f2 x, const f2 y, ObsPair &P)
{
    const float4 a = *reinterpret_cast<const float4 *>(obs);
    const float ox[2] = {a.x, a.z}, oy[2] = {a.y, a.w};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        P.dx[u] = add2(x, bc2(-ox[u]));
        P.dy[u] = add2(y, bc2(-oy[u]));
        P.rr[u] = fma2(P.dy[u], P.dy[u], fma2(P.dx[u], P.dx[u], bc2(1.0f)));
    }
    rcp_block<2, STRICT>(P.rr);
}

__device__ __forceinline__ void obstacle_pair_tail(const ObsPair &P, f2 &sr, f2 &sx, f2 &sy)
{
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        sr = add2(sr, P.rr[u]);
        const f2 r2 = mul2(P.rr[u], P.rr[u]);
        sx = fma2(r2, P.dx[u], sx);
        sy = fma2(r2, P.dy[u], sy);
    }
}

// ---------------------------------------------------------------------------
// Shared obstacle loop of single-warp teams.  A trajectory of T <= 64 samples keeps n_act = ceil(T / 2) lanes busy;
// the other n_help = 32 - n_act lanes of its warp execute every instruction for nothing (T = 50: 7 of 32).  The
// obstacle loop is the one phase whose work per sample is divisible, so with many obstacles it is shared: the owner
// lane sums the obstacles [0, S), a helper lane sums the tail [S, n_obs) of up to k = ceil(n_act / n_help) owners one
// after the other (positions and partial sums travel through the team's operand buffers, which are idle during the cost
// phase), and the owner adds the helper's partial to its own: sr = srA + srB, sx = sxA + sxB, sy = syA + syB - each chain
// from zero.  S = k * Lseg with Lseg = 4 * ceil(n_obs / (4 (k + 1))), so that k owner segments and one tail are equally
// long: T = 50, 256 obstacles: k = 4, S = 208, tail 48 - 208 loop steps per lane instead of 256.  The helper chain is whole
// blocks of four: the last rem = (n_obs - S) mod 4 obstacles belong to the OWNER's chain, after its [0, S) (one piece of
// remainder code per evaluation instead of one per helper segment):
//     chain A = [0, S) ++ [n_obs - rem, n_obs),   chain B = [S, n_obs - rem),   each ascending.  The split is part of the documented summation order (the mirror oracle follows it) whenever
// share_split() > 0: single-warp teams, end-effector cost, at least FGD_SHARE_MIN_OBS obstacles.  Instances without the
// helper code (HELP = false) evaluate both chains in the owner lane - same bits.
// ---------------------------------------------------------------------------
constexpr int FGD_SHARE_MIN_OBS = 64;
__host__ __device__ inline int share_split(int T, int n_obs, bool whole_arm)
{
    if (T > 64 || whole_arm || n_obs < FGD_SHARE_MIN_OBS) return 0;
    const int n_act = (T + 1) / 2, n_help = 32 - n_act;
    if (n_help <= 0) return 0;
    const int k = (n_act + n_help - 1) / n_help;
    const int lseg = 4 * ((n_obs + 4 * (k + 1) - 1) / (4 * (k + 1)));
    const int S = k * lseg;
    return S < n_obs ? S : 0;
}

// obstacles [o0, o1) against the row pair at (x, y), ascending, in blocks of four (o0 and o1 - o0 multiples of four)
template <bool STRICT>
__device__ __forceinline__ void obstacle_range4(const float2 *__restrict__ sObs, int o0, const int o1, const f2 x, const f2 y, f2 &sr, f2 &sx, f2 &sy)
{
#pragma unroll UNROLL_R4
    for (; o0 < o1; o0 += 4) obstacle_block<4, STRICT>(sObs + o0, x, y, sr, sx, sy);
}
// the last obstacles [o0, n_obs) one at a time (any alignment; at most three)
template <bool STRICT>
__device__ __forceinline__ void obstacle_singles(const float2 *__restrict__ sObs, int o0, const int n_obs, const f2 x, const f2 y, f2 &sr, f2 &sx, f2 &sy)
{
#pragma unroll 1
    for (; o0 < n_obs; ++o0) obstacle_block<1, STRICT>(sObs + o0, x, y, sr, sx, sy);
}

// ---------------------------------------------------------------------------
// Cost phase: compute_trajectory_cost + constraintsFulfilled for one trajectory
// whose raw contraction rows are yq (K alpha) and yv (dK alpha).
//   trajectory.py:271-281 (total), :81-88 (max/mean), :183-255 (penalties),
//   :129-137 + robot.py:90-113 (constraint predicates), robot.py:29-36 (fk),
//   environment.py:32-58 (obstacle potential and its (x,y)-gradient).
// The obstacle loop accumulates sum 1/(2 den) and sum d/(2 den)^2; the constant factors
// 0.8 and -0.8 of environment.py:43,57 (times 2 and 4) are applied once per sample.
// ---------------------------------------------------------------------------
// split: share_split() of this launch / obstacle set (0: one chain per sample).  SHARE = 2: the tail chain runs on the
// helper lanes, XA / XB = the team's operand buffers as scratch; 1: both chains in the owner lane; 0: the caller guarantees
// split == 0 (the host launches the helper instance for every scene with an active split) and the code is left out.
// OC > 0: the obstacle count is the compile-time constant OC (the reference's default scene has 11): the loop is straight-line
// code in blocks of 4, 2, 1 - no loop control, and ptxas schedules across the blocks.  Same blocks in the same order.
template <int WPT, bool STRICT, bool ARM, bool PIPE = false, bool OPAQUE_POS = false, int SHARE = 1, int OC = 0>
__device__ __forceinline__ void cost_phase(const DevParams &p, const int T, const float2 *__restrict__ sObs, const int n_obs, const Team<WPT> &G,
                                           const f2 (&yq)[3], const f2 (&yv)[3],
                                           const float *start, const float *goal, float lam_sg, float lam_jl,
                                           Rows<ARM> &Rw, float &loss, float &toc, int &ful,
                                           const int split = 0, float4 *XA = nullptr, float4 *XB = nullptr)
{
    constexpr int NJ = ARM ? 3 : 1;                        // joint positions charged with the obstacle potential
    const int t0 = G.tl * R;
    const int lT = (T - 1) / R, rT = (T - 1) % R;          // team thread / row slot owning sample T-1
    const bool valid0 = t0 < T, valid1 = (t0 + 1) < T;
    float part_p = 0.0f, part_v = 0.0f;
    row_kinematics(p, yq, yv, Rw.q, Rw.v, Rw.sn, Rw.cs);
    f2 px[3], py[3];                                       // fk_1, fk_2, fk_3 = fk      robot.py:33-34, 39-72
    px[0] = mul2(bc2(p.link[0]), Rw.cs[0]); px[1] = fma2(bc2(p.link[1]), Rw.cs[1], px[0]); px[2] = fma2(bc2(p.link[2]), Rw.cs[2], px[1]);
    py[0] = mul2(bc2(p.link[0]), Rw.sn[0]); py[1] = fma2(bc2(p.link[1]), Rw.sn[1], py[0]); py[2] = fma2(bc2(p.link[2]), Rw.sn[2], py[1]);
    // joint-limit penalties and limit predicates of these rows   trajectory.py:215-255, robot.py:104-113
    f2 e3[3], f3[3];
    bool ok0 = true, ok1 = true;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const f2 qb = Rw.q[b], vb = Rw.v[b];
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
    const bool lim_ok = (!valid0 | ok0) & (!valid1 | ok1);
    const f2 es = add2(add2(e3[0], e3[1]), e3[2]), fs = add2(add2(f3[0], f3[1]), f3[2]);
    if (valid0) { part_p = part_p + es.x; part_v = part_v + fs.x; }
    if (valid1) { part_p = part_p + es.y; part_v = part_v + fs.y; }
    // start / goal rows   trajectory.py:183-204  (meaningful in the threads owning rows 0 / T-1)
    float ssp0 = ss3(Rw.q[0].x - start[0], Rw.q[1].x - start[1], Rw.q[2].x - start[2]);
    float ssv0 = ss3(Rw.v[0].x, Rw.v[1].x, Rw.v[2].x);
    float sspT, ssvT;
    if (rT == 0) {
        sspT = ss3(Rw.q[0].x - goal[0], Rw.q[1].x - goal[1], Rw.q[2].x - goal[2]);
        ssvT = ss3(Rw.v[0].x, Rw.v[1].x, Rw.v[2].x);
    } else {
        sspT = ss3(Rw.q[0].y - goal[0], Rw.q[1].y - goal[1], Rw.q[2].y - goal[2]);
        ssvT = ss3(Rw.v[0].y, Rw.v[1].y, Rw.v[2].y);
    }
    // obstacle potential: both samples of this thread (at NJ joint positions) against every obstacle, in blocks of 4.
    // OPAQUE_POS (the LIVE instance): the end-effector position passes through an own-lane shuffle first.  Without it ptxas
    // re-derives px, py from the sines and cosines INSIDE the obstacle loop of that instance (4 extra FFMA2 per block of four
    // obstacles to save four registers; config 4 was 6.7 % slower) - a shuffle result cannot be rematerialised.
    if constexpr (OPAQUE_POS) {
        px[2] = mk2(__shfl_sync(FULL, px[2].x, G.lane), __shfl_sync(FULL, px[2].y, G.lane));
        py[2] = mk2(__shfl_sync(FULL, py[2].x, G.lane), __shfl_sync(FULL, py[2].y, G.lane));
    }
    f2 sr[NJ], sx[NJ], sy[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) { sr[j] = bc2(0.0f); sx[j] = bc2(0.0f); sy[j] = bc2(0.0f); }
    int o = 0;
    if constexpr (OC > 0 && !ARM && WPT == 1) {
#pragma unroll
        for (int ob = 0; ob + 4 <= OC; ob += 4) obstacle_block<4, STRICT>(sObs + ob, px[2], py[2], sr[0], sx[0], sy[0]);
        if constexpr ((OC & 3) >= 2) obstacle_block<2, STRICT>(sObs + (OC & ~3), px[2], py[2], sr[0], sx[0], sy[0]);
        if constexpr (OC & 1) obstacle_block<1, STRICT>(sObs + (OC & ~1), px[2], py[2], sr[0], sx[0], sy[0]);
        o = n_obs;
    } else if (SHARE > 0 && !ARM && WPT == 1 && split > 0) {      // warp-uniform: two chains per sample (see share_split)
        if constexpr (!ARM && WPT == 1 && SHARE == 2) {
            const int n_act = (T + 1) >> 1, n_help = 32 - n_act, k = (n_act + n_help - 1) / n_help, lseg = split / k;
            const int tail_end = n_obs - ((n_obs - split) & 3);
            const bool helper = G.lane >= n_act;
            if (!helper) XA[G.lane] = make_float4(px[2].x, px[2].y, py[2].x, py[2].y);
            __syncwarp();
            f2 x = px[2], y = py[2];
#pragma unroll 1
            for (int j = 0; j < k; ++j) {
                int o0 = j * lseg, o1 = o0 + lseg, own = -1;
                if (helper) {                              // tail of owner `own`, a fresh chain
                    own = (G.lane - n_act) + j * n_help;
                    const float4 pq = XA[own < n_act ? own : 0];
                    x = mk2(pq.x, pq.y); y = mk2(pq.z, pq.w);
                    sr[0] = bc2(0.0f); sx[0] = bc2(0.0f); sy[0] = bc2(0.0f);
                    o0 = split; o1 = own < n_act ? tail_end : split;
                }
                obstacle_range4<STRICT>(sObs, o0, o1, x, y, sr[0], sx[0], sy[0]);
                if (helper && own < n_act) {
                    XA[own] = make_float4(sr[0].x, sr[0].y, sx[0].x, sx[0].y);
                    XB[own] = make_float4(sy[0].x, sy[0].y, 0.0f, 0.0f);
                }
            }
            obstacle_singles<STRICT>(sObs, tail_end, n_obs, px[2], py[2], sr[0], sx[0], sy[0]);      // the owners' chain ends with the remainder
            __syncwarp();
            if (!helper) {
                const float4 r0 = XA[G.lane], r1 = XB[G.lane];
                sr[0] = add2(sr[0], mk2(r0.x, r0.y)); sx[0] = add2(sx[0], mk2(r0.z, r0.w)); sy[0] = add2(sy[0], mk2(r1.x, r1.y));
            }
            __syncwarp();                                  // the scratch rows are free again
        } else {
            f2 tr = bc2(0.0f), tx = bc2(0.0f), ty = bc2(0.0f);
            const int tail_end = n_obs - ((n_obs - split) & 3);
            obstacle_range4<STRICT>(sObs, 0, split, px[2], py[2], sr[0], sx[0], sy[0]);
            obstacle_singles<STRICT>(sObs, tail_end, n_obs, px[2], py[2], sr[0], sx[0], sy[0]);
            obstacle_range4<STRICT>(sObs, split, tail_end, px[2], py[2], tr, tx, ty);
            sr[0] = add2(sr[0], tr); sx[0] = add2(sx[0], tx); sy[0] = add2(sy[0], ty);
        }
        o = n_obs;
    } else if constexpr (!ARM && PIPE && FGD_OBS_PIPE) {
        if (n_obs >= 2) {
            ObsPair A, B;
            obstacle_pair_head<STRICT>(sObs, px[2], py[2], A);
            o = 2;
#pragma unroll UNROLL_PIPE
            for (; o + 4 <= n_obs; o += 4) {              // two pairs per trip so that A / B never have to be copied
                obstacle_pair_head<STRICT>(sObs + o, px[2], py[2], B);
                obstacle_pair_tail(A, sr[0], sx[0], sy[0]);
                obstacle_pair_head<STRICT>(sObs + o + 2, px[2], py[2], A);
                obstacle_pair_tail(B, sr[0], sx[0], sy[0]);
            }
            if (o + 2 <= n_obs) {
                obstacle_pair_head<STRICT>(sObs + o, px[2], py[2], B);
                obstacle_pair_tail(A, sr[0], sx[0], sy[0]);
                obstacle_pair_tail(B, sr[0], sx[0], sy[0]);
                o += 2;
            } else {
                obstacle_pair_tail(A, sr[0], sx[0], sy[0]);
            }
        }
    } else {
#pragma unroll 1
        for (; o + 4 <= n_obs; o += 4) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) obstacle_block<4, STRICT>(sObs + o, px[3 - NJ + j], py[3 - NJ + j], sr[j], sx[j], sy[j]);
        }
        if (o + 2 <= n_obs) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) obstacle_block<2, STRICT>(sObs + o, px[3 - NJ + j], py[3 - NJ + j], sr[j], sx[j], sy[j]);
            o += 2;
        }
    }
    if (o < n_obs) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) obstacle_block<1, STRICT>(sObs + o, px[3 - NJ + j], py[3 - NJ + j], sr[j], sx[j], sy[j]);
    }
    f2 cost = mul2(bc2(1.6f), sr[0]);
    if constexpr (ARM) cost = fma2(bc2(1.6f), sr[2], fma2(bc2(1.6f), sr[1], cost));      // c_1 + c_2 + c_3, explicit fmas as in the oracle
#pragma unroll
    for (int j = 0; j < NJ; ++j) { Rw.gx[j] = mul2(bc2(-3.2f), sx[j]); Rw.gy[j] = mul2(bc2(-3.2f), sy[j]); }
    float part_c = 0.0f, lmax = 0.0f;                 // cost >= 0
    if (valid0) { part_c = part_c + cost.x; lmax = fmaxf(lmax, cost.x); }
    if (valid1) { part_c = part_c + cost.y; lmax = fmaxf(lmax, cost.y); }
    // max / first argmax / sums over t: inside the warp, then across the warps of the team
    float maxc = wmax_nonneg(lmax);                   // lmax >= +0 and never NaN (fmaxf drops a NaN cost): one REDUX instead of a butterfly
    // first sample of this warp that attains the maximum (trajectory.py:97 argmax = first index): two votes instead of a
    // second butterfly; lane l of the warp owns the samples t0w + 2l (slot x) and t0w + 2l + 1 (slot y)
    const unsigned bx = __ballot_sync(FULL, valid0 && cost.x == maxc), by = __ballot_sync(FULL, valid1 && cost.y == maxc);
    const int t0w = t0 - G.lane * R;
    int amax = 0x7fffffff;
    if (by) amax = t0w + (__ffs(by) - 1) * R + 1;
    if (bx) amax = min(amax, t0w + (__ffs(bx) - 1) * R);
    float sum_c = part_c, sum_p = part_p, sum_v = part_v;
    wsum3(sum_c, sum_p, sum_v);
    bool all_ok = __all_sync(FULL, lim_ok);
    if constexpr (WPT == 1) {
        ssp0 = __shfl_sync(FULL, ssp0, 0);
        ssv0 = __shfl_sync(FULL, ssv0, 0);
        sspT = __shfl_sync(FULL, sspT, lT);
        ssvT = __shfl_sync(FULL, ssvT, lT);
    } else {
        float *xc = G.xch + XCH_COST, *xe = G.xch + XCH_ENDS;
        if (G.lane == 0) {
            float *w = xc + G.wit * 8;
            w[0] = maxc; w[1] = __int_as_float(amax); w[2] = sum_c; w[3] = sum_p; w[4] = sum_v; w[5] = all_ok ? 1.0f : 0.0f;
        }
        if (G.tl == 0) { xe[0] = ssp0; xe[1] = ssv0; }
        if (G.tl == lT) { xe[2] = sspT; xe[3] = ssvT; }
        G.sync();
        float pm[WPT], pc[WPT], pp[WPT], pv[WPT];
        all_ok = true;
#pragma unroll
        for (int w = 0; w < WPT; ++w) {
            pm[w] = xc[w * 8]; pc[w] = xc[w * 8 + 2]; pp[w] = xc[w * 8 + 3]; pv[w] = xc[w * 8 + 4];
            all_ok = all_ok && (xc[w * 8 + 5] != 0.0f);
        }
        maxc = pm[0];
#pragma unroll
        for (int w = 1; w < WPT; ++w) maxc = fmaxf(maxc, pm[w]);
        amax = 0x7fffffff;
#pragma unroll
        for (int w = WPT - 1; w >= 0; --w)
            if (pm[w] == maxc) amax = __float_as_int(xc[w * 8 + 1]);      // lowest warp holding the maximum = first argmax
        sum_c = combine_sum<WPT>(pc); sum_p = combine_sum<WPT>(pp); sum_v = combine_sum<WPT>(pv);
        ssp0 = xe[0]; ssv0 = xe[1]; sspT = xe[2]; ssvT = xe[3];
    }
    Rw.amax = amax;
    const float avg = div_T<STRICT>(p, sum_c);
    toc = fmaf(p.lam_max, maxc, p.oml * avg);
    const float sg = (0.5f * ssp0 + 0.5f * sspT) + (0.5f * ssv0 + 0.5f * ssvT);
    const float jl = div_T<STRICT>(p, sum_p) + div_T<STRICT>(p, sum_v);
    loss = fmaf(lam_jl, jl, fmaf(lam_sg, sg, toc));
    // end-point predicates  ||.|| < eps  (robot.py:90-101) without the square roots: the correctly rounded sqrt is monotonic, so
    // sqrt(x) < eps  <=>  x < thr with thr = the smallest float whose square root reaches eps (computed on the host with the IEEE
    // sqrt, fgd_api.cu sqrt_threshold) - exact in both math modes (the oracle takes the square roots)
    const bool ends_ok = (ssp0 < p.thr_pos) && (sspT < p.thr_pos) && (ssv0 < p.thr_vel) && (ssvT < p.thr_vel);
    ful = (ends_ok && all_ok) ? 1 : 0;
}

// ---------------------------------------------------------------------------
// Gradient phase: rows of  G_q = toc_g + lam_sg*sgp_g + lam_jl*jp_g  and
// G_v = lam_sg*sgv_g + lam_jl*jv_g  (trajectory.py:289-295, :91-126, robot.py:75-87),
// written as the operands of the backward contraction: XA = G_q, XB = -G_v
// (dK^T = -dK bit-exactly, checked in fgd_create()).  nz[w][r] collects the rows whose
// velocity gradient is not identically zero (consumed by contract_back); for WPT > 1
// the warp masks go through the team scratch and are read back after the team
// barrier that precedes the backward contraction (load_nz).
// ---------------------------------------------------------------------------
template <int WPT, bool ARM>
__device__ __forceinline__ void grad_phase(const DevParams &p, const int T, const Team<WPT> &G, const Rows<ARM> &Rw, const float *start, const float *goal,
                                           float lam_sg, float lam_jl, float4 *XA, float4 *XB, unsigned (&nz)[WPT][R])
{
    constexpr int NJ = ARM ? 3 : 1;
    const int ta = G.tl * R, tb = ta + 1;
    const float w_hi = p.lam_max + p.w_avg;
    const f2 wt = mk2((ta == Rw.amax) ? w_hi : p.w_avg, (tb == Rw.amax) ? w_hi : p.w_avg);
    const f2 cgx = mul2(wt, Rw.gx[NJ - 1]), cgy = mul2(wt, Rw.gy[NJ - 1]);
    f2 xs[3], ys[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { xs[k] = neg2(mul2(bc2(p.link[k]), Rw.sn[k])); ys[k] = mul2(bc2(p.link[k]), Rw.cs[k]); }
    const f2 Sx = add2(add2(xs[0], xs[1]), xs[2]), Sy = add2(add2(ys[0], ys[1]), ys[2]);
    const f2 Cx[3] = {xs[0], add2(xs[0], xs[1]), add2(add2(xs[0], xs[1]), xs[2])};
    const f2 Cy[3] = {ys[0], add2(ys[0], ys[1]), add2(add2(ys[0], ys[1]), ys[2])};
    f2 gq[3], gv[3];
    const bool a0 = (ta == 0), aT = (ta == T - 1), bT = (tb == T - 1);     // tb >= 1
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const f2 Jx = sub2(add2(xs[k], Sx), Cx[k]);
        const f2 Jy = sub2(add2(ys[k], Sy), Cy[k]);
        f2 tg = fma2(cgy, Jy, mul2(cgx, Jx));
        if constexpr (ARM) {                                   // + joints 2 and 1: J_j[k] = sum_{m = k..j}
            if (k <= 1) {
                const f2 J2x = (k == 0) ? add2(xs[0], xs[1]) : xs[1], J2y = (k == 0) ? add2(ys[0], ys[1]) : ys[1];
                tg = fma2(mul2(wt, Rw.gy[1]), J2y, fma2(mul2(wt, Rw.gx[1]), J2x, tg));
            }
            if (k == 0) tg = fma2(mul2(wt, Rw.gy[0]), ys[0], fma2(mul2(wt, Rw.gx[0]), xs[0], tg));
        }
        const f2 qk = Rw.q[k], vk = Rw.v[k];
        f2 sgp, sgv;
        sgp.x = a0 ? (qk.x - start[k]) : (aT ? (qk.x - goal[k]) : 0.0f);
        sgp.y = bT ? (qk.y - goal[k]) : 0.0f;
        sgv.x = (a0 || aT) ? vk.x : 0.0f;
        sgv.y = bT ? vk.y : 0.0f;
        const bool m0 = p.cvdl ? (qk.x > p.q_hi || qk.x < p.q_lo) : true;
        const bool m1 = p.cvdl ? (qk.y > p.q_hi || qk.y < p.q_lo) : true;
        const f2 jpg = sel2(m0, m1, mul2(mul2(add2(qk, bc2(-p.mean_q)), bc2(p.inv_std2)), bc2(p.inv_T)), bc2(0.0f));
        const bool n0 = p.cvdl ? (fabsf(vk.x) > p.v_hi) : true;
        const bool n1 = p.cvdl ? (fabsf(vk.y) > p.v_hi) : true;
        const f2 jvg = sel2(n0, n1, mul2(mul2(vk, bc2(p.inv_vmax2)), bc2(p.inv_T)), bc2(0.0f));
        gq[k] = fma2(bc2(lam_jl), jpg, fma2(bc2(lam_sg), sgp, tg));
        gv[k] = fma2(bc2(lam_jl), jvg, mul2(bc2(lam_sg), sgv));
    }
    const bool wa = ta < T, wb = tb < T;
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
    if constexpr (WPT == 1) {
        nz[0][0] = ma; nz[0][1] = mb;
    } else if (G.lane == 0) {
        unsigned *xn = reinterpret_cast<unsigned *>(G.xch + XCH_NZ);
        xn[G.wit * R] = ma; xn[G.wit * R + 1] = mb;
    }
}

// WPT > 1: the non-zero-row masks of all warps of the team (call after the team barrier that follows grad_phase)
template <int WPT>
__device__ __forceinline__ void load_nz(const Team<WPT> &G, unsigned (&nz)[WPT][R])
{
    if constexpr (WPT > 1) {
        const unsigned *xn = reinterpret_cast<const unsigned *>(G.xch + XCH_NZ);
#pragma unroll
        for (int w = 0; w < WPT; ++w) { nz[w][0] = xn[w * R]; nz[w][1] = xn[w * R + 1]; }
    }
}

// alpha-gradient rows from the backward contraction: (K^T G_q + dK^T G_v) J^T
__device__ __forceinline__ void backward_rows(const DevParams &p, const f2 (&y1)[3], const f2 (&y2)[3], f2 (&g)[3])
{
    const f2 r0 = add2(y1[0], y2[0]), r1 = add2(y1[1], y2[1]), r2 = add2(y1[2], y2[2]);
#pragma unroll
    for (int b = 0; b < 3; ++b)
        g[b] = fma2(r2, bc2(p.J[b * 3 + 2]), fma2(r1, bc2(p.J[b * 3 + 1]), mul2(r0, bc2(p.J[b * 3]))));
}

}  // namespace fgd
