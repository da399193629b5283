// Separable Gaussian, generic tile kernel (any dtype, any channel count, any
// width): ONE launch does both passes -- the halo'd input tile is staged in
// shared memory (zero outside the image = mode "constant", cval 0), the row pass
// writes a shared intermediate, the column pass writes global.  No intermediate
// image, no host sync between passes, weights travel as kernel parameters.
//
// Replaces g_gaussian_{row,col}_{one,four}_channel + the host choreography of
// _gaussian_greyscale/_gaussian_rgba (src/millipyde_image.cpp:146-381, :725-856):
// two launches with a hipStreamSynchronize between them, a scratch image, and
// weights pushed through a mutable __constant__ symbol shared by all workers.
//
// This is the correctness-first path and the fallback for ragged widths; the
// fp32 roofline path is kernels/gaussian_stream.cuh.
#pragma once
#include "common.cuh"

namespace mpk {



template <typename T>
__device__ __forceinline__ T fma_t(T a, T b, T c);
template <>
__device__ __forceinline__ float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <>
__device__ __forceinline__ double fma_t<double>(double a, double b, double c) { return fma(a, b, c); }

// CLAMP0: fmax(0, .) after the column pass (the reference's fp64 kernel, :240).
template <typename T, int C, bool CLAMP0>
__global__ void __launch_bounds__(256)
gauss_tile_kernel(const T *__restrict__ in, T *__restrict__ out, int width, int height, int tile_h,
                  int tile_wp, const __grid_constant__ GaussParams<T> gp)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int R = gp.radius;
    const int in_w = (tile_wp + 2 * R) * C;  // elements per staged row
    const int in_h = tile_h + 2 * R;
    const int out_w = tile_wp * C;
    T *s_in = reinterpret_cast<T *>(smem_raw);
    T *s_h = s_in + (size_t)in_h * in_w;

    const int x0 = blockIdx.x * tile_wp, y0 = blockIdx.y * tile_h;
    const int row_elems = width * C;
    const int gx_base = (x0 - R) * C;

    for (int i = threadIdx.x; i < in_h * in_w; i += blockDim.x) {
        int r = i / in_w, j = i - r * in_w;
        int gy = y0 - R + r, gx = gx_base + j;
        T v = (T)0;
        if (gy >= 0 && gy < height && gx >= 0 && gx < row_elems) v = in[(size_t)gy * row_elems + gx];
        s_in[i] = v;
    }
    __syncthreads();

    // row pass, taps in the reference's order k = -R .. R
    for (int i = threadIdx.x; i < in_h * out_w; i += blockDim.x) {
        int r = i / out_w, j = i - r * out_w;
        const T *p = s_in + r * in_w + j + R * C;
        T acc = (T)0;
        for (int k = -R; k <= R; ++k) acc = fma_t<T>(p[k * C], gp.w[k < 0 ? -k : k], acc);
        s_h[i] = acc;
    }
    __syncthreads();

    for (int i = threadIdx.x; i < tile_h * out_w; i += blockDim.x) {
        int r = i / out_w, j = i - r * out_w;
        int gy = y0 + r, gx = x0 * C + j;
        if (gy >= height || gx >= row_elems) continue;
        const T *p = s_h + (r + R) * out_w + j;
        T acc = (T)0;
        for (int k = -R; k <= R; ++k) acc = fma_t<T>(p[k * out_w], gp.w[k < 0 ? -k : k], acc);
        if (CLAMP0) acc = acc > (T)0 ? acc : (T)0;
        out[(size_t)gy * row_elems + gx] = acc;
    }
}

// ---------------------------------------------------------------- fp64 greyscale, fixed radius
// The reference's greyscale layout (rgb2grey of a uint8 image is fp64).  Same rule and the same
// accumulation order as gauss_tile_kernel<double, 1, CLAMP0> -- every output adds its taps in the
// order k = -R .. R with fma -- so results are bit-identical to it; what changes is the shape of the
// work: the radius is a template parameter (R = 8: the reference's 17 taps, R = 16: the oracle's 33
// taps at sigma = 2), every thread keeps a run of outputs in registers and streams the inputs past
// them (one shared load feeds up to 8 DFMAs), and both passes read shared memory without conflicts.
//
// Tile: 32 columns x TH rows with TH + 2R = 128 staged rows.  Pass 1: task = (staged row, run of 8
// columns), lanes on consecutive rows; row pitches of 2R + 34 and 34 doubles put consecutive lanes 16
// bytes apart.  Pass 2: task = (column, run of P2 rows), lanes on consecutive columns, so shared
// loads and global stores are contiguous.
template <int R>
struct GaussF64Geom {
    static constexpr int TW = 32, IN_H = 128, TH = IN_H - 2 * R;
    static constexpr int P1 = 8, P2 = TH / 16;          // 512 tasks per pass = 2 per thread
    static_assert(TH % 16 == 0 && P2 >= 1, "pass-2 runs");
    static constexpr int IN_W = TW + 2 * R;
    static constexpr int PITCH_IN = IN_W + 2, PITCH_H = TW + 2;
    static constexpr size_t SMEM = ((size_t)IN_H * PITCH_IN + (size_t)IN_H * PITCH_H) * sizeof(double);
};

template <int R, bool CLAMP0>
__global__ void __launch_bounds__(256, 2)
gauss_f64_kernel(const double *__restrict__ in, double *__restrict__ out, int width, int height,
                 const __grid_constant__ GaussParams<double> gp, int tma)
{
    using G = GaussF64Geom<R>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_in = reinterpret_cast<double *>(smem_raw);   // [128][PITCH_IN]
    double *s_h = s_in + G::IN_H * G::PITCH_IN;              // [128][PITCH_H]
    __shared__ uint64_t s_bar;
    const int x0 = blockIdx.x * G::TW, y0 = blockIdx.y * G::TH;
    const int tid = threadIdx.x;

    if (tma) {
        // ---- staging by the TMA unit (rows of whole 16-byte vectors: even width, aligned base): the
        // in-image part of every staged row is one 1-D bulk copy, issued by lane 0 of each warp for rows
        // w, w + 8, ...; what lies outside the image is zeroed (mode "constant", cval 0).  No load
        // instructions, no registers in flight, and the other CTA of the SM computes meanwhile.
        const int lane = tid & 31, w = __shfl_sync(0xffffffffu, tid >> 5, 0);
        if (tid == 0) {
            mbar_init(&s_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        const int gx_lo = x0 - R, gy_lo = y0 - R;
        const int c_lo = gx_lo < 0 ? -gx_lo : 0, c_hi = G::IN_W < width - gx_lo ? G::IN_W : width - gx_lo;
        const int r_lo = gy_lo < 0 ? -gy_lo : 0, r_hi = G::IN_H < height - gy_lo ? G::IN_H : height - gy_lo;
        if (lane == 0) {
            const uint32_t row_bytes = (uint32_t)(c_hi - c_lo) * 8u;
            if (w == 0) mbar_expect_tx(&s_bar, row_bytes * (uint32_t)(r_hi - r_lo));
            const uint32_t bar32 = smem_addr(&s_bar);
            uint32_t sdst = smem_addr(s_in) + (uint32_t)((r_lo + w) * G::PITCH_IN + c_lo) * 8u;
            const double *gsrc = in + ((size_t)(gy_lo + r_lo + w) * width + gx_lo + c_lo);
            for (int r = r_lo + w; r < r_hi; r += 8, gsrc += (size_t)8 * width, sdst += 8u * G::PITCH_IN * 8u)
                bulk_g2s_raw(sdst, gsrc, row_bytes, bar32);
        }
        if (c_lo > 0 || c_hi < G::IN_W || r_lo > 0 || r_hi < G::IN_H) {
            for (int r = w; r < G::IN_H; r += 8) {
                double *row = s_in + r * G::PITCH_IN;
                if (r < r_lo || r >= r_hi) {
                    for (int c = lane; c < G::IN_W; c += 32) row[c] = 0.0;
                } else {
                    for (int c = lane; c < c_lo; c += 32) row[c] = 0.0;
                    for (int c = c_hi + lane; c < G::IN_W; c += 32) row[c] = 0.0;
                }
            }
        }
        if (w == 0) mbar_wait_suspend(&s_bar, 0, 2000);
    } else {
        // ---- staging, 8 loads in flight per thread; zero outside the image (mode "constant", cval 0)
        constexpr int N_IN = G::IN_H * G::IN_W;
        static_assert(N_IN % (256 * 8) == 0, "whole batches");
        for (int base = 0; base < N_IN; base += 256 * 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + tid + 256 * u;
                const int r = idx / G::IN_W, j = idx - r * G::IN_W;
                const int gy = y0 - R + r, gx = x0 - R + j;
                v[u] = 0.0;
                if (gy >= 0 && gy < height && gx >= 0 && gx < width) v[u] = __ldg(in + (size_t)gy * width + gx);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + tid + 256 * u;
                const int r = idx / G::IN_W, j = idx - r * G::IN_W;
                s_in[r * G::PITCH_IN + j] = v[u];
            }
        }
    }
    __syncthreads();

    // ---- pass 1 (rows)
#pragma unroll 1
    for (int u = 0; u < 2; ++u) {
        const int task = tid + 256 * u;
        const int row = task & (G::IN_H - 1), run = task >> 7;
        const double *src = s_in + row * G::PITCH_IN + G::P1 * run;
        double acc[G::P1];
#pragma unroll
        for (int i = 0; i < G::P1; ++i) acc[i] = 0.0;
#pragma unroll
        for (int j2 = 0; j2 < (G::P1 + 2 * R) / 2; ++j2) {
            const double2 v2 = *reinterpret_cast<const double2 *>(src + 2 * j2);
            const double v[2] = {v2.x, v2.y};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = 2 * j2 + h;
#pragma unroll
                for (int i = 0; i < G::P1; ++i) {
                    const int k = j - i - R;
                    if (k >= -R && k <= R) acc[i] = fma(v[h], gp.w[k < 0 ? -k : k], acc[i]);
                }
            }
        }
        double *dst = s_h + row * G::PITCH_H + G::P1 * run;
#pragma unroll
        for (int i = 0; i < G::P1; i += 2) *reinterpret_cast<double2 *>(dst + i) = make_double2(acc[i], acc[i + 1]);
    }
    __syncthreads();

    // ---- pass 2 (columns)
#pragma unroll 1
    for (int u = 0; u < 2; ++u) {
        const int task = tid + 256 * u;
        const int col = task & 31, run = task >> 5;
        const double *src = s_h + (G::P2 * run) * G::PITCH_H + col;
        double acc[G::P2];
#pragma unroll
        for (int i = 0; i < G::P2; ++i) acc[i] = 0.0;
#pragma unroll
        for (int j = 0; j < G::P2 + 2 * R; ++j) {
            const double v = src[j * G::PITCH_H];
#pragma unroll
            for (int i = 0; i < G::P2; ++i) {
                const int k = j - i - R;
                if (k >= -R && k <= R) acc[i] = fma(v, gp.w[k < 0 ? -k : k], acc[i]);
            }
        }
        const int gx = x0 + col;
#pragma unroll
        for (int i = 0; i < G::P2; ++i) {
            const int gy = y0 + G::P2 * run + i;
            if (gx < width && gy < height) {
                double a = acc[i];
                if (CLAMP0) a = a > 0.0 ? a : 0.0;
                out[(size_t)gy * width + gx] = a;
            }
        }
    }
}

// Packed RGBA8, the reference's integer rule (src/millipyde_image.cpp:247-381):
// per byte lane sum_k (int)(byte * w_k) -- each product truncated before the
// integer add -- then & 0xff; the column pass forces bits 24..31 to 0xff.
// Weights are the reference's doubles (float expf, normalised in double).
__global__ void __launch_bounds__(256)
gauss_rgba8_tile_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int width,
                        int height, int tile_h, int tile_w, const __grid_constant__ GaussParams<double> gp)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int R = gp.radius;
    const int in_w = tile_w + 2 * R, in_h = tile_h + 2 * R;
    uint32_t *s_in = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *s_h = s_in + (size_t)in_h * in_w;
    const int x0 = blockIdx.x * tile_w, y0 = blockIdx.y * tile_h;

    for (int i = threadIdx.x; i < in_h * in_w; i += blockDim.x) {
        int r = i / in_w, j = i - r * in_w;
        int gy = y0 - R + r, gx = x0 - R + j;
        uint32_t v = 0;
        if (gy >= 0 && gy < height && gx >= 0 && gx < width) v = in[(size_t)gy * width + gx];
        s_in[i] = v;
    }
    __syncthreads();

    for (int i = threadIdx.x; i < in_h * tile_w; i += blockDim.x) {
        int r = i / tile_w, j = i - r * tile_w;
        const uint32_t *p = s_in + r * in_w + j + R;
        int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        for (int k = -R; k <= R; ++k) {
            uint32_t v = p[k];
            double w = gp.w[k < 0 ? -k : k];
            s0 += (int)((v & 0xff) * w);
            s1 += (int)(((v >> 8) & 0xff) * w);
            s2 += (int)(((v >> 16) & 0xff) * w);
            s3 += (int)(((v >> 24) & 0xff) * w);
        }
        s_h[i] = ((uint32_t)(s3 & 0xff) << 24) | ((uint32_t)(s2 & 0xff) << 16) |
                 ((uint32_t)(s1 & 0xff) << 8) | (uint32_t)(s0 & 0xff);
    }
    __syncthreads();

    for (int i = threadIdx.x; i < tile_h * tile_w; i += blockDim.x) {
        int r = i / tile_w, j = i - r * tile_w;
        int gy = y0 + r, gx = x0 + j;
        if (gy >= height || gx >= width) continue;
        const uint32_t *p = s_h + (r + R) * tile_w + j;
        int s0 = 0, s1 = 0, s2 = 0;
        for (int k = -R; k <= R; ++k) {
            uint32_t v = p[k * tile_w];
            double w = gp.w[k < 0 ? -k : k];
            s0 += (int)((v & 0xff) * w);
            s1 += (int)(((v >> 8) & 0xff) * w);
            s2 += (int)(((v >> 16) & 0xff) * w);
        }
        out[(size_t)gy * width + gx] = 0xff000000u | ((uint32_t)(s2 & 0xff) << 16) |
                                       ((uint32_t)(s1 & 0xff) << 8) | (uint32_t)(s0 & 0xff);
    }
}

// Integer form of the same rule.  (int)(b * w) for a byte b and a double weight w equals
// umulhi(b, M) with M = floor(w * 2^32) (or that + 1) unless b * w lies within ~255 * 2^-32 of an
// integer; the host checks all 256 bytes x 9 weights against the double products for the sigma at
// hand and only then launches this kernel, so the result is bit-identical by construction while the
// inner loop is three integer instructions per byte lane instead of I2F.F64 + DMUL + F2I.F64 (the
// conversions run on the quarter-rate pipe and made the kernel 1.6 % of HBM).
struct GaussU8Params {
    int radius;
    uint32_t m[kGaussMaxRadius + 1];
};

__global__ void __launch_bounds__(256)
gauss_rgba8_int_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int width, int height,
                       int tile_h, int tile_w, const __grid_constant__ GaussU8Params gp)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int R = gp.radius;
    const int in_w = tile_w + 2 * R, in_h = tile_h + 2 * R;
    uint32_t *s_in = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *s_h = s_in + (size_t)in_h * in_w;
    const int x0 = blockIdx.x * tile_w, y0 = blockIdx.y * tile_h;

    for (int i = threadIdx.x; i < in_h * in_w; i += blockDim.x) {
        int r = i / in_w, j = i - r * in_w;
        int gy = y0 - R + r, gx = x0 - R + j;
        uint32_t v = 0;
        if (gy >= 0 && gy < height && gx >= 0 && gx < width) v = __ldg(in + (size_t)gy * width + gx);
        s_in[i] = v;
    }
    __syncthreads();

    for (int i = threadIdx.x; i < in_h * tile_w; i += blockDim.x) {
        int r = i / tile_w, j = i - r * tile_w;
        const uint32_t *p = s_in + r * in_w + j + R;
        uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        for (int k = -R; k <= R; ++k) {
            const uint32_t v = p[k], m = gp.m[k < 0 ? -k : k];
            s0 += __umulhi(v & 0xff, m);
            s1 += __umulhi((v >> 8) & 0xff, m);
            s2 += __umulhi((v >> 16) & 0xff, m);
            s3 += __umulhi(v >> 24, m);
        }
        s_h[i] = ((s3 & 0xff) << 24) | ((s2 & 0xff) << 16) | ((s1 & 0xff) << 8) | (s0 & 0xff);
    }
    __syncthreads();

    for (int i = threadIdx.x; i < tile_h * tile_w; i += blockDim.x) {
        int r = i / tile_w, j = i - r * tile_w;
        int gy = y0 + r, gx = x0 + j;
        if (gy >= height || gx >= width) continue;
        const uint32_t *p = s_h + (r + R) * tile_w + j;
        uint32_t s0 = 0, s1 = 0, s2 = 0;
        for (int k = -R; k <= R; ++k) {
            const uint32_t v = p[k * tile_w], m = gp.m[k < 0 ? -k : k];
            s0 += __umulhi(v & 0xff, m);
            s1 += __umulhi((v >> 8) & 0xff, m);
            s2 += __umulhi((v >> 16) & 0xff, m);
        }
        out[(size_t)gy * width + gx] = 0xff000000u | ((s2 & 0xff) << 16) | ((s1 & 0xff) << 8) | (s0 & 0xff);
    }
}

// ---------------------------------------------------------------- RGBA8, fp32 chain form
// The same rule -- sum over the taps of (int)(byte * w), horizontally, then vertically on the
// horizontal result, alpha := 0xff -- with ONE instruction per tap and channel pair.  A running sum
// t = 2^23 + n (n the integer accumulated so far) is an fp32 whose ulp is 1, so
//     t' = fma.rm(b, w, t)                    (round toward -inf)
// is exactly 2^23 + n + floor(b * w): the product is exact inside the FMA, t is an integer, and the
// round-down lands on the integer below.  The host verifies floor(b * (float)w) == (int)(b * w)
// for all 256 bytes x 9 weights of the sigma at hand before launching (else the integer kernel
// above runs).  Packed as fma.rm.f32x2 over the (r, g) and (b, -) halves of a pixel, a tap costs two
// issue slots per pixel instead of twelve -- (r, g) packed, b as a scalar FFMA.RM (half the pipe time of a packed
// pair whose other lane would be the alpha the rule drops).
//
// Effective radius.  A tap with 255 * w < 1 contributes (int)(byte * w) = 0 for every byte, in both
// passes (the vertical pass reads bytes again), so it need not be evaluated: at sigma = 2 the
// reference's 17 taps are 11 (w[6] = 0.0022 -> 255 w = 0.57).  The launcher picks the smallest
// instantiated radius R that holds every tap with floor(255 * w) >= 1 (bit-exact by construction).
//
// Tile: 32 x 48 output pixels per CTA.  Staging converts the (32 + 2R) x (48 + 2R) input pixels to
// float4 once (alpha dropped).  Pass 1: thread = (input row, run of 8 pixels), lanes on consecutive
// rows -- the row pitch of 33 + 2R float4 (odd) spreads them over the banks.  Pass 2: thread =
// (column, run of 6 rows), lanes on consecutive columns, so shared loads and the 4-byte global
// stores are contiguous.
constexpr int kU8R = 8;               // the reference's radius (17 taps), src/millipyde_image.cpp:744
constexpr int kU8TW = 32, kU8TH = 48;
template <int R>
struct U8Geom {
    static constexpr int InW = kU8TW + 2 * R, InH = kU8TH + 2 * R;    // 48 x 64 at R = 8
    static constexpr int PitchIn = InW + 1, PitchH = kU8TW + 1;        // in float4 units
    static constexpr size_t Smem = ((size_t)InH * PitchIn + (size_t)InH * PitchH) * 16;
    static_assert(InH <= 64, "pass 1 has 64 thread rows");
};

struct GaussU8ChainParams {
    unsigned long long ww[kU8R + 1];  // ((float)w[d], (float)w[d]) packed
};

__device__ __forceinline__ uint64_t ffma2_rm(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t d;
    asm("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

__device__ __forceinline__ float ffma_rm(float a, float b, float c)
{
    float d;
    asm("fma.rm.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

template <int R>
__global__ void __launch_bounds__(256, 2)
gauss_rgba8_chain_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int width, int height,
                         const __grid_constant__ GaussU8ChainParams gp)
{
    using G = U8Geom<R>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_in = reinterpret_cast<float4 *>(smem_raw);                 // [InH][InW + 1]
    float4 *s_h = s_in + G::InH * G::PitchIn;                            // [InH][33]
    const int x0 = blockIdx.x * kU8TW, y0 = blockIdx.y * kU8TH;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr float kM = 8388608.f;  // 2^23
    const uint64_t m2 = ((uint64_t)__float_as_uint(kM) << 32) | __float_as_uint(kM);

    // ---- staging: InH rows x InW pixels, row-major over the threads (a row is contiguous bytes).
    // All loads are issued before the first conversion.
    {
        constexpr int N = G::InH * G::InW, PER = (N + 255) / 256;
        uint32_t v[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int idx = tid + 256 * u;
            const int r = idx / G::InW, j = idx - r * G::InW;
            const int gy = y0 - R + r, gx = x0 - R + j;
            v[u] = 0;
            if (idx < N && gy >= 0 && gy < height && gx >= 0 && gx < width) v[u] = __ldg(in + (size_t)gy * width + gx);
        }
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int idx = tid + 256 * u;
            const int r = idx / G::InW, j = idx - r * G::InW;
            if (idx < N)
                s_in[r * G::PitchIn + j] = make_float4((float)(v[u] & 0xff), (float)((v[u] >> 8) & 0xff),
                                                       (float)((v[u] >> 16) & 0xff), 0.f);
        }
    }
    __syncthreads();

    // ---- pass 1 (rows): thread = (row, run of 8 output pixels)
    {
        const int row = (warp & 1) * 32 + lane, run = warp >> 1;
        if (row < G::InH) {
            const float4 *src = s_in + row * G::PitchIn + 8 * run;
            uint64_t a_lo[8];   // (r, g)
            float a_b[8];       // b: a scalar FFMA.RM -- the packed form would spend half its lanes on the dropped alpha
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a_lo[i] = m2;
                a_b[i] = kM;
            }
#pragma unroll
            for (int j = 0; j < 8 + 2 * R; ++j) {
                const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(src + j);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int k = j - i - R;  // tap index of input j for output i
                    if (k >= -R && k <= R) {
                        const uint64_t w = gp.ww[k < 0 ? -k : k];
                        a_lo[i] = ffma2_rm(v.x, w, a_lo[i]);
                        a_b[i] = ffma_rm(__uint_as_float((uint32_t)v.y), __uint_as_float((uint32_t)w), a_b[i]);
                    }
                }
            }
            float4 *dst = s_h + row * G::PitchH + 8 * run;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float r = __uint_as_float((uint32_t)a_lo[i]), g = __uint_as_float((uint32_t)(a_lo[i] >> 32));
                dst[i] = make_float4(r - kM, g - kM, a_b[i] - kM, 0.f);  // exact: integers below 256
            }
        }
    }
    __syncthreads();

    // ---- pass 2 (columns): thread = (column, run of 6 output rows)
    {
        const int col = lane, run = warp;
        const float4 *src = s_h + (6 * run) * G::PitchH + col;
        uint64_t a_lo[6];
        float a_b[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            a_lo[i] = m2;
            a_b[i] = kM;
        }
#pragma unroll
        for (int j = 0; j < 6 + 2 * R; ++j) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(src + j * G::PitchH);
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const int k = j - i - R;
                if (k >= -R && k <= R) {
                    const uint64_t w = gp.ww[k < 0 ? -k : k];
                    a_lo[i] = ffma2_rm(v.x, w, a_lo[i]);
                    a_b[i] = ffma_rm(__uint_as_float((uint32_t)v.y), __uint_as_float((uint32_t)w), a_b[i]);
                }
            }
        }
        const int gx = x0 + col;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int gy = y0 + 6 * run + i;
            if (gx < width && gy < height) {
                // t = 2^23 + n: the low mantissa byte IS n
                const uint32_t r = (uint32_t)a_lo[i] & 0xff, g = (uint32_t)(a_lo[i] >> 32) & 0xff;
                const uint32_t b = __float_as_uint(a_b[i]) & 0xff;
                out[(size_t)gy * width + gx] = 0xff000000u | (b << 16) | (g << 8) | r;
            }
        }
    }
}

// (Measured and not kept: the same chain with the three live channels of two pixels packed into
// three register pairs -- no alpha lane, 1.5 instead of 2 FFMA2 per pixel and tap -- on 64 x 48 tiles,
// also as a persistent kernel prefetching the next tile: 52-55 us per 4K image against 54 us for the
// kernel above.  ncu: FMA pipe 44 %, issue 53 %, LSU 52 % -- neither form is bound by the FMA count
// any more; the tile phases (stage, barrier, rows, barrier, columns) at 16-24 warps per SM are.
// Also measured and not kept: the packed pixels staged by the TMA unit into a raw buffer borrowed
// from s_h and converted from there -- 60.5 us: the extra shared-memory round trip and barrier cost
// more than the global-load phase they replace.  The fp64 kernel above, which needs no conversion,
// gains 19 % from the same staging.)

}  // namespace mpk
