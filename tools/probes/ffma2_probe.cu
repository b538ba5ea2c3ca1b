// Issue-slot probe for packed fp32 (FFMA2) on sm_100a: does one FFMA2 cost one issue slot for two FMAs?
// Each variant runs the same number of fp32 FMAs per thread; "mixed" variants interleave ALU-pipe integer
// instructions so that the issue port, not the FMA pipe, is the bound.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float ffma1(float a, float b, float c)
{
    float r;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b)
{
    unsigned r;
    asm volatile("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}

template <int kVariant> __global__ void probe(float* out, int iters, float seed)
{
    float s[16];
    unsigned long long p[8];
    unsigned u[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = seed + i + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float2 v = make_float2(s[2 * i], s[2 * i + 1]);
        p[i] = *reinterpret_cast<unsigned long long*>(&v);
        u[i] = threadIdx.x * 7 + i;
    }
    const float m = seed * 0.5f, c = seed * 0.25f;
    float2 mm = make_float2(m, m), cc = make_float2(c, c);
    const unsigned long long m2 = *reinterpret_cast<unsigned long long*>(&mm), c2 = *reinterpret_cast<unsigned long long*>(&cc);
    for (int it = 0; it < iters; ++it) {
        if (kVariant == 0 || kVariant == 2) {  // 16 scalar FMAs
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i] = ffma1(s[i], m, c);
        } else {  // 8 packed FMAs
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], m2, c2);
        }
        if (kVariant >= 2) {  // + 8 ALU-pipe instructions
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = lop(u[i], u[(i + 1) & 7]);
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += s[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float2 v = *reinterpret_cast<float2*>(&p[i]);
        acc += v.x + v.y + float(u[i]);
    }
    if (acc == 12345.678f) out[0] = acc;
}

template <int kVariant> static float run(int iters)
{
    float* d;
    cudaMalloc(&d, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<kVariant><<<148 * 4, 256>>>(d, iters, 1.0f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<kVariant><<<148 * 4, 256>>>(d, iters, 1.0f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaFree(d);
    return ms;
}

int main()
{
    const int iters = 20000;
    const double warps_per_smsp = 4.0 * 256 / 32 / 4;  // 8
    const float t0 = run<0>(iters), t1 = run<1>(iters), t2 = run<2>(iters), t3 = run<3>(iters);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    auto cyc = [&](float ms, double insts) { return ms * 1e-3 * khz * 1e3 / (iters * warps_per_smsp * insts); };
    printf("clock attr %d kHz (nominal; cycles below assume it)\n", khz);
    printf("v0 16xFFMA            %.3f ms  %.3f cyc/warp-inst (16 inst/iter)\n", t0, cyc(t0, 16));
    printf("v1  8xFFMA2           %.3f ms  %.3f cyc/warp-inst ( 8 inst/iter)\n", t1, cyc(t1, 8));
    printf("v2 16xFFMA  + 8xLOP3  %.3f ms  %.3f cyc/warp-inst (24 inst/iter)\n", t2, cyc(t2, 24));
    printf("v3  8xFFMA2 + 8xLOP3  %.3f ms  %.3f cyc/warp-inst (16 inst/iter)\n", t3, cyc(t3, 16));
    printf("ratios: v1/v0 %.3f   v3/v2 %.3f\n", t1 / t0, t3 / t2);
    return 0;
}
