// Pointwise kernels: brightness / adjust_gamma / colorize / rgb2grey on the three
// layout families.  All are one-touch HBM streams: 128-bit loads and stores,
// grid-stride over a grid sized in multiples of the SM count, no shared memory
// (except the 768-byte LUT of the RGBA8 path).
//
// Replaces g_color_to_greyscale, g_brightness_*, g_adjust_gamma_*,
// g_colorize_four_channel (src/millipyde_image.cpp:50-66, :384-524), which use
// one scalar element per thread on a (W/32, H/32) x (32, 32) grid.
#pragma once
#include "common.cuh"

namespace mpk {

// ---------------------------------------------------------------- fp32, HWC
// Every warp-level access is 32 consecutive 16-byte vectors (512 contiguous bytes); a warp owns
// 96 consecutive vectors per iteration (= 384 floats, a multiple of 3 and 4, so the channel of
// every element is known from the lane: vector v starts at channel (4v) mod C).  Measured on
// B200 (tools/stream_peak.cu): this shape reaches 93 % of the copy bandwidth, a thread-contiguous
// 48-byte chunk only 75 %.
//
// TAB = true: every image of a batched launch has its own program (Generator streams with
// random_* parameters); the records live in device memory and a block copies its image's record
// into shared memory once.
template <int C, bool TAB = false>
__global__ void __launch_bounds__(256)
pw_f32_kernel(const float *__restrict__ in, float *__restrict__ out, size_t n,
              const __grid_constant__ PwProgram prog_one, const float *const *__restrict__ in_tab = nullptr,
              float *const *__restrict__ out_tab = nullptr, const PwProgram *__restrict__ prog_tab = nullptr)
{
    if (in_tab) {  // batched launch: blockIdx.y selects the image
        in = in_tab[blockIdx.y];
        out = out_tab[blockIdx.y];
    }
    __shared__ PwProgram s_prog;
    if (TAB) {
        if (threadIdx.x < sizeof(PwProgram) / 4)
            reinterpret_cast<int *>(&s_prog)[threadIdx.x] = reinterpret_cast<const int *>(prog_tab + blockIdx.y)[threadIdx.x];
        __syncthreads();
    }
    const PwProgram &prog = TAB ? s_prog : prog_one;
    const size_t nvec = n / 4;
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const float4 *in4 = reinterpret_cast<const float4 *>(in);
    float4 *out4 = reinterpret_cast<float4 *>(out);
    for (size_t base = warp * 96; base < nvec; base += nwarps * 96) {
        float4 v[3];
        bool ok[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const size_t idx = base + lane + 32 * j;
            ok[j] = idx < nvec;
            if (ok[j]) v[j] = ld_stream(in4 + idx);
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            // base is a multiple of 96, so vector idx starts at channel (4 * idx) mod 3 == (lane + 2j) mod 3
            const int phase = (C == 3) ? (lane + 2 * j) % 3 : 0;
            float(&r)[4] = *reinterpret_cast<float(*)[4]>(&v[j]);
            pw_apply_tile<C, 4>(prog, r, phase);
        }
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (ok[j]) st_stream(out4 + base + lane + 32 * j, v[j]);
    }
    // the last n % 4 floats
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const size_t e = nvec * 4 + threadIdx.x;
        out[e] = pw_apply<C>(prog, in[e], (int)(e % C));
    }
}

// rgb2grey on fp32 H x W x {3,4} -> H x W, with optional pointwise programs
// before (per colour channel) and after (on the grey value) for fused chains.
// C = 3: a warp streams 96 consecutive vectors (128 pixels) with coalesced 16-byte loads into
// its own 1.5 KB of shared memory, each lane then reads ITS four pixels back as three LDS.128
// (lane pitch 48 B = 3 x 16 B: conflict-free) and the warp stores 32 consecutive float4.
// C = 4: a vector is a pixel, so loads are coalesced as they are; greys leave as 4-byte stores.
// TAB = true: per-image (pre, post) program pairs from device memory, as in pw_f32_kernel.
template <int C, bool TAB = false>
__global__ void __launch_bounds__(256)
grey_f32_kernel(const float *__restrict__ in, float *__restrict__ out, size_t npix,
                const __grid_constant__ PwProgram pre_one, const __grid_constant__ PwProgram post_one,
                const float *const *__restrict__ in_tab = nullptr, float *const *__restrict__ out_tab = nullptr,
                const PwProgram *__restrict__ prog_tab = nullptr)  // [image][2] = pre, post
{
    static_assert(C == 3 || C == 4, "colour input");
    if (in_tab) {
        in = in_tab[blockIdx.y];
        out = out_tab[blockIdx.y];
    }
    __shared__ PwProgram s_prog[2];
    if (TAB) {
        if (threadIdx.x < 2 * sizeof(PwProgram) / 4)
            reinterpret_cast<int *>(s_prog)[threadIdx.x] =
                reinterpret_cast<const int *>(prog_tab + 2 * (size_t)blockIdx.y)[threadIdx.x];
        __syncthreads();
    }
    const PwProgram &pre = TAB ? s_prog[0] : pre_one;
    const PwProgram &post = TAB ? s_prog[1] : post_one;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    if (C == 3) {
        __shared__ __align__(16) float stage[8][384];
        const size_t ngroups = npix / 128;  // 128 pixels = 96 vectors in, 32 vectors out
        const float4 *in4 = reinterpret_cast<const float4 *>(in);
        for (size_t g = warp; g < ngroups; g += nwarps) {
#pragma unroll
            for (int j = 0; j < 3; ++j)
                *reinterpret_cast<float4 *>(&stage[wib][4 * (lane + 32 * j)]) = ld_stream(in4 + g * 96 + lane + 32 * j);
            __syncwarp();
            float r[12];
#pragma unroll
            for (int j = 0; j < 3; ++j)
                *reinterpret_cast<float4 *>(&r[4 * j]) = *reinterpret_cast<const float4 *>(&stage[wib][12 * lane + 4 * j]);
            __syncwarp();
            pw_apply_tile<3, 12>(pre, r, 0);
            float o[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) o[p] = luma_f32(r[3 * p], r[3 * p + 1], r[3 * p + 2]);
            pw_apply_tile<1, 4>(post, o, 0);
            st_stream(reinterpret_cast<float4 *>(out) + g * 32 + lane, make_float4(o[0], o[1], o[2], o[3]));
        }
        // ragged tail (< 128 pixels)
        const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (size_t p = ngroups * 128 + gtid; p < npix; p += (size_t)gridDim.x * blockDim.x) {
            float cr = pw_apply<C>(pre, in[p * C + 0], 0);
            float cg = pw_apply<C>(pre, in[p * C + 1], 1);
            float cb = pw_apply<C>(pre, in[p * C + 2], 2);
            out[p] = pw_apply<1>(post, luma_f32(cr, cg, cb), 0);
        }
    } else {
        const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (size_t p = gtid; p < npix; p += (size_t)gridDim.x * blockDim.x) {
            float4 v = ld_stream(reinterpret_cast<const float4 *>(in) + p);
            float r[4] = {v.x, v.y, v.z, v.w};
            pw_apply_tile<4, 4>(pre, r, 0);
            float o[1] = {luma_f32(r[0], r[1], r[2])};
            pw_apply_tile<1, 1>(post, o, 0);
            out[p] = o[0];
        }
    }
}

// ------------------------------------------------ uint8 colour -> fp64 grey
// The reference's only channel-generic kernel (src/millipyde_image.cpp:50-66):
// bytes 0..2 of each pixel, alpha ignored, result double.
__device__ __forceinline__ double luma_u8(unsigned r, unsigned g, unsigned b)
{
    return fmin(1.0, (0.2125 * r + 0.7154 * g + 0.0721 * b) / 255);
}

__global__ void __launch_bounds__(256)
grey_rgba8_kernel(const uint32_t *__restrict__ in, double *__restrict__ out, size_t npix)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t ngroups = npix / 4;
    for (size_t g = tid; g < ngroups; g += stride) {
        uint4 px = ld_stream(reinterpret_cast<const uint4 *>(in) + g);
        const uint32_t w[4] = {px.x, px.y, px.z, px.w};
        double o[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) o[p] = luma_u8(w[p] & 0xff, (w[p] >> 8) & 0xff, (w[p] >> 16) & 0xff);
        double2 *dst = reinterpret_cast<double2 *>(out + g * 4);
        st_stream(dst, make_double2(o[0], o[1]));
        st_stream(dst + 1, make_double2(o[2], o[3]));
    }
    for (size_t p = ngroups * 4 + tid; p < npix; p += stride) {
        uint32_t w = in[p];
        out[p] = luma_u8(w & 0xff, (w >> 8) & 0xff, (w >> 16) & 0xff);
    }
}

// Generic channel count (RGB8 and anything the reference's kernel would accept).
__global__ void __launch_bounds__(256)
grey_u8_generic_kernel(const uint8_t *__restrict__ in, double *__restrict__ out, size_t npix, int channels)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += stride) {
        const uint8_t *q = in + p * channels;
        out[p] = luma_u8(q[0], q[1], q[2]);
    }
}

// fp64 H x W x 3 -> H x W (skimage.color.rgb2gray on a float64 image).
__global__ void __launch_bounds__(256)
grey_f64_kernel(const double *__restrict__ in, double *__restrict__ out, size_t npix, int channels)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += stride) {
        const double *q = in + p * channels;
        out[p] = fmin(1.0, 0.2125 * q[0] + 0.7154 * q[1] + 0.0721 * q[2]);
    }
}

// ---------------------------------------------------------------- fp64, H x W

// v ** g in double for the oracle rule (numpy's float64 power; the tests allow 1e-12).  The library
// pow is ~150 fp64 instructions per sample and B200's fp64 pipe is narrow (a 4K image took 79 us, a
// quarter of the HBM rate), so the exponents that are a few multiplications or a square root are
// evaluated as such (correctly rounded steps: within 2 ulp of pow), and any other exponent as
// exp(g * log v) for v > 0 (relative error <= |g log v| * 2^-52: 1e-14 on image data); only samples
// that are not positive and finite take the library call, which knows every special case.
__device__ __forceinline__ double pow64(double v, double g)
{
    if (g == 2.0) return v * v;
    if (g == 1.0) return v;
    if (g == 0.5 && v >= 0.0) return sqrt(v);
    if (g == 1.5 && v >= 0.0) return v * sqrt(v);
    if (g == 3.0) return v * v * v;
    if (g == 4.0) {
        const double q = v * v;
        return q * q;
    }
    if (v > 0.0 && v < 1.7e308) return exp(g * log(v));
    return pow(v, g);
}

// `float_pow`: reference semantics -- powf on float-converted operands
// (src/millipyde_image.cpp:447); otherwise pow in double (the test oracle).
__device__ __forceinline__ double pw64(const PwOp64 &op, double v, bool float_pow)
{
    if (op.kind == PW_BRIGHTNESS) {
        v = v + op.a;
    } else if (op.kind == PW_GAMMA) {
        v = float_pow ? op.b * powf(v, op.a) : op.b * pow64(v, op.a);
    } else {
        return v;
    }
    v = (v < 0 ? 0 : v);
    v = (v > 1 ? 1 : v);
    return v;
}

__global__ void __launch_bounds__(256)
pw_f64_kernel(const double *__restrict__ in, double *__restrict__ out, size_t n, PwOp64 op, bool float_pow)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nvec = n / 2;
    for (size_t i = tid; i < nvec; i += stride) {
        double2 v = ld_stream(reinterpret_cast<const double2 *>(in) + i);
        v.x = pw64(op, v.x, float_pow);
        v.y = pw64(op, v.y, float_pow);
        st_stream(reinterpret_cast<double2 *>(out) + i, v);
    }
    if (tid == 0 && (n & 1)) out[n - 1] = pw64(op, in[n - 1], float_pow);
}

// ------------------------------------------------------------ packed RGBA8
// brightness, adjust_gamma and colorize are all byte -> byte maps per colour
// channel (alpha untouched), so a chain of them is three 256-entry tables.  The
// tables are built on the device with exactly the reference's expressions
// (device powf, float product, truncating casts), then every pixel is three
// shared-memory lookups: bit-identical to evaluating the ops one by one.


__device__ __forceinline__ unsigned u8_apply(const U8Op &op, unsigned v, int ch)
{
    switch (op.kind) {
        case PW_BRIGHTNESS: {
            int t = (int)v + op.d8;
            return (unsigned)(op.d8 > 0 ? min(255, t) : max(0, t)) & 0xffu;
        }
        case PW_GAMMA: {
            double t = op.b * (255 * powf((double)v / 255, op.a));
            t = (t < 0 ? 0 : t);
            return t > 255 ? 255u : (unsigned)(unsigned char)t;
        }
        case PW_COLORIZE: {
            double t = (double)v * (ch == 0 ? op.a : (ch == 1 ? op.b : op.c));
            return t > 255 ? 255u : (unsigned)(unsigned char)t;
        }
        default: return v;
    }
}

// The composed byte tables of a program: lut[ch][v] = ops applied in order to byte v of channel ch.
// Built once per distinct program (the launcher caches the 768-byte result per device), so the
// double-precision pow of the gamma rule is paid 768 times in total, not 768 times per CTA.
__global__ void __launch_bounds__(256)
u8_lut_kernel(uint8_t *__restrict__ lut, const __grid_constant__ U8Program prog)
{
    const int ch = blockIdx.x;
    unsigned v = threadIdx.x;
    for (int i = 0; i < prog.n; ++i) v = u8_apply(prog.ops[i], v, ch);
    lut[ch * 256 + threadIdx.x] = (uint8_t)v;
}

// Four 16-byte vectors per thread, 256 apart, so every warp access is 512 contiguous bytes and a
// CTA streams 16 KB; many small CTAs keep more bytes in flight than a short persistent grid.
constexpr int kU8PwVecs = 4;

__global__ void __launch_bounds__(256)
pw_rgba8_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, size_t npix,
                const uint8_t *__restrict__ lut_g)
{
    __shared__ __align__(16) uint8_t lut[3][256];
    if (threadIdx.x < 192)
        reinterpret_cast<uint32_t *>(&lut[0][0])[threadIdx.x] = __ldg(reinterpret_cast<const uint32_t *>(lut_g) + threadIdx.x);
    __syncthreads();
    auto map = [&](uint32_t w) -> uint32_t {
        return (w & 0xff000000u) | ((uint32_t)lut[2][(w >> 16) & 0xff] << 16) |
               ((uint32_t)lut[1][(w >> 8) & 0xff] << 8) | lut[0][w & 0xff];
    };
    const size_t ngroups = npix / 4;
    const size_t base = (size_t)blockIdx.x * (256 * kU8PwVecs) + threadIdx.x;
    uint4 px[kU8PwVecs];
#pragma unroll
    for (int u = 0; u < kU8PwVecs; ++u)
        if (base + 256 * u < ngroups) px[u] = ld_stream(reinterpret_cast<const uint4 *>(in) + base + 256 * u);
#pragma unroll
    for (int u = 0; u < kU8PwVecs; ++u)
        if (base + 256 * u < ngroups) {
            px[u].x = map(px[u].x);
            px[u].y = map(px[u].y);
            px[u].z = map(px[u].z);
            px[u].w = map(px[u].w);
            st_stream(reinterpret_cast<uint4 *>(out) + base + 256 * u, px[u]);
        }
    if (blockIdx.x == 0 && threadIdx.x < (npix & 3)) {  // the last npix % 4 pixels
        const size_t p = ngroups * 4 + threadIdx.x;
        out[p] = map(in[p]);
    }
}

}  // namespace mpk
