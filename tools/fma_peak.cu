// Microbenchmark: fp32 FMA issue rate on this GPU for the three instruction forms the Gaussian can use.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_peak fma_peak.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct W { float w[8]; unsigned long long ww[8]; };

__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c)
{ uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, const __grid_constant__ W p, float seed)
{
    float x = seed + threadIdx.x * 1e-3f;
    float a[16];
    uint64_t b[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = x + i; b[i] = ((uint64_t)__float_as_uint(x + i) << 32) | __float_as_uint(x - i); }
    uint64_t xv = ((uint64_t)__float_as_uint(x) << 32) | __float_as_uint(x * 0.5f);
    float y = x * 0.25f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (MODE == 0) a[i] = fmaf(p.w[r], x, a[i]);          // FFMA R, R, UR/const, R
                else if (MODE == 1) a[i] = fmaf(y, x, a[i]);          // FFMA R, R, R, R
                else b[i] = ffma2(p.ww[r], xv, b[i]);                 // FFMA2 with uniform operand
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) { s += a[i]; s += __uint_as_float((uint32_t)b[i]) + __uint_as_float((uint32_t)(b[i] >> 32)); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int fma_per_instr)
{
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    float *out; cudaMalloc(&out, sms * 8 * 256 * 4);
    W p; for (int i = 0; i < 8; ++i) { p.w[i] = 1e-6f * i; unsigned u; memcpy(&u, &p.w[i], 4); p.ww[i] = ((unsigned long long)u << 32) | u; }
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms * 8, 256>>>(out, 100, p, 1.f);
    cudaEventRecord(e0);
    k<MODE><<<sms * 8, 256>>>(out, iters, p, 1.f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double instr = (double)sms * 8 * 256 * iters * 8 * 16;
    double fma = instr * fma_per_instr;
    printf("%-28s %8.3f ms  %7.2f TFMA/s  %6.1f FMA/clk/SM (at %.0f MHz max clock)\n", name, ms, fma / ms / 1e9,
           fma / (ms * 1e-3) / sms / (khz * 1e3), khz / 1e3);
    cudaFree(out);
}

int main()
{
    run<0>("FFMA  reg,const/UR,reg", 1);
    run<1>("FFMA  reg,reg,reg", 1);
    run<2>("FFMA2 reg,UR,reg", 2);
    return 0;
}
