// How many warps does FFMA2 need to saturate the fp32 pipe?  (1 CTA per SM, varying warps.)
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
struct W { unsigned long long ww[12]; };
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c)
{ uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int NACC>
__global__ void k(float *out, int iters, const __grid_constant__ W p, float seed)
{
    float x = seed + threadIdx.x * 1e-3f;
    uint64_t b[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) b[i] = ((uint64_t)__float_as_uint(x + i) << 32) | __float_as_uint(x - i);
    uint64_t xv = ((uint64_t)__float_as_uint(x) << 32) | __float_as_uint(x * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < NACC; ++i) b[i] = ffma2(p.ww[(i + r) % 12], xv, b[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += __uint_as_float((uint32_t)b[i]) + __uint_as_float((uint32_t)(b[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
void run(int threads)
{
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *out; cudaMalloc(&out, sms * 1024 * 4);
    W p; for (int i = 0; i < 12; ++i) { float w = 1e-6f * i; unsigned u; memcpy(&u, &w, 4); p.ww[i] = ((unsigned long long)u << 32) | u; }
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NACC><<<sms, threads>>>(out, 100, p, 1.f);
    cudaEventRecord(e0);
    k<NACC><<<sms, threads>>>(out, iters, p, 1.f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)sms * threads * iters * 4 * NACC * 2;
    printf("acc=%2d warps/SM=%2d  %7.3f ms  %6.1f FMA/clk/SM\n", NACC, threads / 32, ms, fma / (ms * 1e-3) / sms / (khz * 1e3));
    cudaFree(out);
}
int main()
{
    for (int t : {128, 256, 384, 512, 640, 768, 1024}) run<23>(t);
    for (int t : {256, 640}) run<8>(t);
    return 0;
}
