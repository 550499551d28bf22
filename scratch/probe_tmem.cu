// Probe: per-lane K-table reads from TMEM (tcgen05.ld.32x32b.x4) vs shared memory (LDS.128)
// inside a contraction-shaped loop (per k: 16 B own-lane table + 16 B broadcast operand + 6 FFMA2).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef float2 f2;
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 bc2(float a) { return make_float2(a, a); }
constexpr int T = 50, NW = 8, COLS = 256;

template <int MODE>   // 0 = smem table, 1 = TMEM table
__global__ void __launch_bounds__(NW * 32, 2) k(int reps, const float *tab, float *out)
{
    extern __shared__ __align__(16) float sm[];
    float *sK = sm;                         // [T][32][4]
    float4 *sX = reinterpret_cast<float4 *>(sm + T * 128);   // [NW][T]
    __shared__ uint32_t tmem_base_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < T * 128; i += NW * 32) sK[i] = tab[i];
    for (int i = threadIdx.x; i < NW * T; i += NW * 32) sX[i] = make_float4(0.001f * i, 0.002f, 0.003f, 0.f);
    uint32_t tbase = 0;
    if (MODE >= 1) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"l"((uint64_t)__cvta_generic_to_shared(&tmem_base_s)), "n"(COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
        tbase = tmem_base_s;
        if (warp < 4) {          // each lane quadrant gets its own copy of the table
            const uint32_t ta = tbase + ((uint32_t)(warp * 32) << 16);
            for (int kk = 0; kk < T; ++kk) {
                const float4 v = *reinterpret_cast<const float4 *>(tab + (kk * 32 + lane) * 4);
                asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ta + 4 * kk), "r"(__float_as_uint(v.x)),
                             "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)));
            }
            asm volatile("tcgen05.wait::st.sync.aligned;");
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
    } else {
        __syncthreads();
    }
    const uint32_t ta = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    const float4 *X = sX + warp * T;
    const float *kd = sK + lane * 4;
    f2 y[6];
    for (int i = 0; i < 6; ++i) y[i] = bc2(0.f);
    if (MODE == 2 || MODE == 3) {
        // software pipelined: loads of block i+1 are in flight while block i is consumed
        uint32_t nx[5][4];
        auto issue = [&](int k0) {
#pragma unroll
            for (int u = 0; u < 5; ++u)
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(nx[u][0]), "=r"(nx[u][1]), "=r"(nx[u][2]), "=r"(nx[u][3]) : "r"(ta + 4 * (k0 + u)));
        };
        issue(0);
        for (int r = 0; r < reps; ++r) {
#pragma unroll 1
            for (int k0 = 0; k0 < T; k0 += 5) {
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float4 kv[5];
#pragma unroll
                for (int u = 0; u < 5; ++u) kv[u] = make_float4(__uint_as_float(nx[u][0]), __uint_as_float(nx[u][1]), __uint_as_float(nx[u][2]), __uint_as_float(nx[u][3]));
                issue(k0 + 5 < T ? k0 + 5 : 0);
                if (MODE == 3) {
#pragma unroll
                    for (int u = 0; u < 5; ++u) { y[0].x += kv[u].x; y[1].x += kv[u].y; y[2].x += kv[u].z; y[3].x += kv[u].w; }
                } else {
#pragma unroll
                    for (int u = 0; u < 5; ++u) {
                        const float4 x = X[k0 + u];
                        const f2 kk = make_float2(kv[u].x, kv[u].y), dk = make_float2(kv[u].z, kv[u].w);
                        y[0] = fma2(kk, bc2(x.x), y[0]); y[1] = fma2(kk, bc2(x.y), y[1]); y[2] = fma2(kk, bc2(x.z), y[2]);
                        y[3] = fma2(dk, bc2(x.x), y[3]); y[4] = fma2(dk, bc2(x.y), y[4]); y[5] = fma2(dk, bc2(x.z), y[5]);
                    }
                }
            }
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    } else
    for (int r = 0; r < reps; ++r) {
#pragma unroll 1
        for (int k0 = 0; k0 < T; k0 += 5) {
            float4 kv[5];
            if (MODE == 1) {
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    uint32_t a, b, c, d;
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(ta + 4 * (k0 + u)));
                    kv[u] = make_float4(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c), __uint_as_float(d));
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else if (MODE == 0) {
#pragma unroll
                for (int u = 0; u < 5; ++u) kv[u] = *reinterpret_cast<const float4 *>(kd + (k0 + u) * 128);
            } else {
#pragma unroll
                for (int u = 0; u < 5; ++u) kv[u] = make_float4(y[0].x, y[1].y, y[2].x, y[3].y);
            }
#pragma unroll
            for (int u = 0; u < 5; ++u) {
                const float4 x = X[k0 + u];
                const f2 kk = make_float2(kv[u].x, kv[u].y), dk = make_float2(kv[u].z, kv[u].w);
                y[0] = fma2(kk, bc2(x.x), y[0]); y[1] = fma2(kk, bc2(x.y), y[1]); y[2] = fma2(kk, bc2(x.z), y[2]);
                y[3] = fma2(dk, bc2(x.x), y[3]); y[4] = fma2(dk, bc2(x.y), y[4]); y[5] = fma2(dk, bc2(x.z), y[5]);
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 6; ++i) s += y[i].x + y[i].y;
    out[blockIdx.x * NW * 32 + threadIdx.x] = s;
    if (MODE >= 1) {
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(COLS));
    }
}

template <int MODE> void run(const char *name, const float *tab, float *out, float *h)
{
    const int grid = 148 * 2, reps = 2000;
    const size_t smem = T * 128 * 4 + NW * T * 16;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int it = 0; it < 4; ++it) {
        cudaEventRecord(e0); k<MODE><<<grid, NW * 32, smem>>>(reps, tab, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (it && ms < best) best = ms;
    }
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, 64 * 4, cudaMemcpyDeviceToHost);
    const double ksteps = (double)grid * NW * reps * T;
    printf("%-26s %8.3f ms  %.3f SM-cycles per warp-k-step (16 warps/SM)  checksum %.6e  err=%s\n", name, best,
           best * 1e-3 * 1.965e9 * 148 / ksteps, (double)h[5], cudaGetErrorString(e));
}
int main()
{
    float *tab, *out, h[64];
    cudaMalloc(&tab, T * 128 * 4); cudaMalloc(&out, 148 * 2 * NW * 32 * 4);
    float *ht = new float[T * 128];
    for (int i = 0; i < T * 128; ++i) ht[i] = 1e-3f * (i % 977);
    cudaMemcpy(tab, ht, T * 128 * 4, cudaMemcpyHostToDevice);
    run<0>("smem", tab, out, h);
    run<1>("tmem", tab, out, h);
    run<2>("tmem-pipelined", tab, out, h);
    run<3>("tmem-loads-only", tab, out, h);
    run<4>("no-table (x+FFMA2 only)", tab, out, h);
    return 0;
}
