// What can a streaming elementwise kernel reach on this GPU?  Variants of out[i] = f(in[i]) over 199 MB
// (one 4K RGB fp32 image), cycling over 8 buffer pairs so L2 cannot help.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int VPT>
__global__ void __launch_bounds__(256) k(const float4 *__restrict__ in, float4 *__restrict__ out, size_t n4, float d)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i * VPT < n4; i += stride) {
        float4 v[VPT];
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            size_t idx = (MODE & 4) ? (i + j * (n4 / VPT)) : (i * VPT + j);   // bit 2: strided planes vs contiguous chunk
            if (MODE & 1) v[j] = __ldcs(in + idx); else v[j] = in[idx];
        }
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            v[j].x = fminf(fmaxf(v[j].x + d, 0.f), 1.f); v[j].y = fminf(fmaxf(v[j].y + d, 0.f), 1.f);
            v[j].z = fminf(fmaxf(v[j].z + d, 0.f), 1.f); v[j].w = fminf(fmaxf(v[j].w + d, 0.f), 1.f);
        }
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            size_t idx = (MODE & 4) ? (i + j * (n4 / VPT)) : (i * VPT + j);
            if (MODE & 2) __stcs(out + idx, v[j]); else out[idx] = v[j];
        }
    }
}
template <int MODE, int VPT>
void run(const char *name, float4 **a, float4 **b, size_t n4, int blocks)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int r = 0; r < 4; ++r) k<MODE, VPT><<<blocks, 256>>>(a[r], b[r], n4, 0.1f);
    cudaEventRecord(e0);
    const int reps = 40;
    for (int r = 0; r < reps; ++r) k<MODE, VPT><<<blocks, 256>>>(a[r % 8], b[r % 8], n4, 0.1f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-34s blocks=%6d  %7.2f us  %7.1f GB/s\n", name, blocks, ms / reps * 1e3, 2.0 * n4 * 16 / (ms / reps) / 1e6);
}
int main()
{
    const size_t n4 = (size_t)2160 * 3840 * 3 / 4;
    float4 *a[8], *b[8];
    for (int i = 0; i < 8; ++i) { cudaMalloc(&a[i], n4 * 16); cudaMalloc(&b[i], n4 * 16); cudaMemset(a[i], 0, n4 * 16); }
    int sms = 148;
    for (int bl : {sms * 4, sms * 8, sms * 16, sms * 32}) {
        run<0, 1>("plain ld/st, 1 vec", a, b, n4, bl);
        run<3, 1>("cs ld/st, 1 vec", a, b, n4, bl);
        run<3, 3>("cs, 3 vec contiguous (current)", a, b, n4, bl);
        run<7, 3>("cs, 3 vec strided planes", a, b, n4, bl);
        run<7, 4>("cs, 4 vec strided planes", a, b, n4, bl);
        run<4, 4>("plain, 4 vec strided planes", a, b, n4, bl);
    }
    int full = (int)((n4 + 255) / 256);
    run<0, 1>("plain, one vec per thread (no loop)", a, b, n4, full);
    run<3, 1>("cs, one vec per thread (no loop)", a, b, n4, full);
    return 0;
}
