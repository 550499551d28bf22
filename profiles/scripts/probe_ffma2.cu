// Throughput probe: FFMA vs FFMA2 (fma.rn.f32x2) and co-issue with ALU / LDS work on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(int iters, float seed, float *sink, int *isink)
{
    __shared__ float4 sm[256];
    sm[threadIdx.x] = make_float4(seed, seed, seed, seed);
    __syncthreads();
    float2 a[8];
    for (int i = 0; i < 8; ++i) a[i] = make_float2(seed + threadIdx.x + i, seed + i);
    const float2 m = make_float2(0.999f, 0.998f), c = make_float2(1e-3f, 2e-3f);
    int x0 = threadIdx.x, x1 = 3, x2 = 5, x3 = 7;
    float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0 || MODE == 2) {          // scalar FFMA: 16 per u
#pragma unroll
                for (int i = 0; i < 8; ++i) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
            } else {                               // FFMA2: 8 per u (same flops)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(a[i], m, c);
            }
            if (MODE == 2 || MODE == 3) {          // + 8 ALU ops per u
                x0 = (x0 ^ x1) + x2; x1 = (x1 ^ x2) + x3; x2 = (x2 ^ x3) + x0; x3 = (x3 ^ x0) + x1;
            }
            if (MODE == 4) {                       // FFMA2 + 2 LDS.128 per u
                float4 v = sm[(threadIdx.x + u) & 255]; float4 w = sm[(u * 7 + it) & 255];
                acc.x += v.x; acc.y += w.y;
            }
        }
    }
    float r = 0;
    for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y;
    r += acc.x + acc.y;
    if (r == 123456.789f) sink[0] = r;
    if ((x0 ^ x1 ^ x2 ^ x3) == 0x12345678) isink[0] = x0;
}
template <int MODE> void run(const char *name)
{
    float *sink; int *isink; cudaMalloc(&sink, 4); cudaMalloc(&isink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, iters = 2048;
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0); k<MODE><<<grid, 256>>>(iters, 1.0f + rep, sink, isink); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = (double)grid * 256 * iters * 8 * 16 * 2 / (ms * 1e-3) * 1e-12;
        if (rep && fl > best) best = fl;
    }
    printf("%-28s %.2f TFLOP/s (FMA flops only)\n", name, best);
}
int main()
{
    run<0>("FFMA"); run<1>("FFMA2"); run<2>("FFMA + ALU(8 per 16)"); run<3>("FFMA2 + ALU(8 per 8)"); run<4>("FFMA2 + 2 LDS.128 per 8");
    return 0;
}
