// Index / resampling kernels: transpose, fliplr, rotate.
//
// Replaces g_transpose<T>, g_flip_horizontal<T>, g_rotate<T>
// (src/millipyde_image.cpp:73-139).  The pixel is the unit of motion; a pixel
// is K 32-bit words (K = 1: RGBA8 or fp32 grey, 2: fp64 grey, 3: fp32 RGB,
// 4: fp32 RGBA), so one kernel template serves every layout bit-exactly.
#pragma once
#include "common.cuh"
#include "gather_params.cuh"

namespace mpk {

// ------------------------------------------------------------------ transpose
// 32 x 32 pixel tile through shared memory.  A tile row is 32*K words plus K
// words of padding, i.e. a pitch of 33*K == K (mod 32): the transposed
// read-back of word j = y*K + c of output row x hits bank (j + const) mod 32,
// conflict-free for every K.  (The reference tile is unpadded: 32-way
// conflicts, src/millipyde_image.cpp:76,94.)  Global accesses are contiguous
// 32*K-word row segments on both sides.
template <int K>
__global__ void __launch_bounds__(256)
transpose_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int width, int height)
{
    constexpr int P = 33 * K;
    __shared__ uint32_t tile[32 * P];
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    const int tw = min(32, width - x0), th = min(32, height - y0);

    // load: row r of the tile = words [x0*K, (x0+tw)*K) of input row y0+r
    for (int i = threadIdx.x; i < 32 * 32 * K; i += 256) {
        int r = i / (32 * K), j = i - r * (32 * K);
        if (r < th && j < tw * K)
            tile[r * P + j] = in[((size_t)(y0 + r) * width + x0) * K + j];
    }
    __syncthreads();
    // store: output row x0+r holds pixels y0..y0+th-1 -> words j = y*K + c
    for (int i = threadIdx.x; i < 32 * 32 * K; i += 256) {
        int r = i / (32 * K), j = i - r * (32 * K);
        int y = j / K, c = j - y * K;
        if (r < tw && y < th)
            out[((size_t)(x0 + r) * height + y0) * K + j] = tile[y * P + r * K + c];
    }
}

// TMA-fed variant (needs (W*K) % 4 == 0 and (H*K) % 4 == 0, i.e. 16-byte row pitches on both
// sides).  The 32 rows of a tile arrive as 32 one-dimensional bulk copies (cp.async.bulk, SASS
// UBLKCP) issued by the 32 lanes of warp 0 -- no load instructions, no shared-memory stores, no
// registers in flight -- landing at a pitch of 32*K + 4 words (16-byte aligned, and the 4-word
// skew spreads a column over the banks).  Each thread then gathers four consecutive words of an
// OUTPUT row (4 x LDS.32) and writes them as one 16-byte store, so the store side moves 128-bit
// vectors too.  Many small CTAs (12.8 KB of shared memory each for RGB fp32) overlap each other's
// copy latency.  Images of a batch are addressed through pointer tables (blockIdx.z).
template <int K>
__global__ void __launch_bounds__(256)
transpose_tma_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int width, int height,
                     const uint32_t *const *__restrict__ in_tab = nullptr,
                     uint32_t *const *__restrict__ out_tab = nullptr)
{
    constexpr int P = 32 * K + 4;
    __shared__ __align__(128) uint32_t tile[32 * P];
    __shared__ __align__(8) uint64_t bar;
    if (in_tab) {
        in = in_tab[blockIdx.z];
        out = out_tab[blockIdx.z];
    }
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    const int tw = min(32, width - x0), th = min(32, height - y0);
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid < 32) {
        const uint32_t row_bytes = (uint32_t)tw * K * 4u;
        if (tid == 0) mbar_expect_tx(&bar, row_bytes * (uint32_t)th);
        __syncwarp();
        if (tid < th)
            bulk_g2s(tile + tid * P, in + ((size_t)(y0 + tid) * width + x0) * K, row_bytes, &bar);
    }
    mbar_wait(&bar, 0);
    // output row (x0 + r) holds pixels y0 .. y0+th-1: th*K words, written as th*K/4 vectors
    const int vec_per_row = th * K / 4;
    for (int i = tid; i < tw * vec_per_row; i += 256) {
        const int r = i / vec_per_row, q = i - r * vec_per_row;
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = 4 * q + e;  // word of the output row: pixel y = j / K, channel c = j % K
            const int y = j / K, c = j - y * K;
            w[e] = tile[y * P + r * K + c];
        }
        uint4 *dst = reinterpret_cast<uint4 *>(out + ((size_t)(x0 + r) * height + y0) * K) + q;
        st_stream(dst, make_uint4(w[0], w[1], w[2], w[3]));
    }
}

// --------------------------------------------------------------------- fliplr
// Vector path (width % 4 == 0): a thread moves 4 pixels = K 16-byte vectors,
// reading the mirrored 4-pixel block and reversing the pixel order in
// registers; both sides are aligned 128-bit accesses.
template <int K>
__global__ void __launch_bounds__(256)
fliplr_vec_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, int width, size_t nblocks4)
{
    const int bpr = width / 4;  // 4-pixel blocks per row
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < nblocks4; b += stride) {
        size_t row = b / bpr;
        int bx = (int)(b - row * bpr);
        const uint4 *src = in + (row * bpr + (bpr - 1 - bx)) * K;
        uint4 v[K];
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = ld_stream(src + k);
        const uint32_t *w = reinterpret_cast<const uint32_t *>(v);
        uint4 o[K];
        uint32_t *ow = reinterpret_cast<uint32_t *>(o);
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int c = 0; c < K; ++c) ow[p * K + c] = w[(3 - p) * K + c];
        uint4 *dst = out + (row * bpr + bx) * K;
#pragma unroll
        for (int k = 0; k < K; ++k) st_stream(dst + k, o[k]);
    }
}

// Warp-cooperative path (width % 128 == 0): a warp mirrors 128 pixels = 32*K vectors.  Both the
// global loads and the global stores are 32 consecutive 16-byte vectors per instruction; the
// pixel reversal happens on the way through a 32*K-vector shared-memory stage (the per-lane
// 16*K-byte chunks of fliplr_vec_kernel cost ~25 % of the bandwidth, tools/stream_peak.cu).
template <int K>
__global__ void __launch_bounds__(256)
fliplr_warp_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, int width, size_t nunits,
                   const uint4 *const *__restrict__ in_tab = nullptr, uint4 *const *__restrict__ out_tab = nullptr)
{
    if (in_tab) {
        in = in_tab[blockIdx.y];
        out = out_tab[blockIdx.y];
    }
    __shared__ uint4 stage[8][32 * K];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const int upr = width / 128;  // 128-pixel units per row
    for (size_t u = warp; u < nunits; u += nwarps) {
        const size_t row = u / upr;
        const int ux = (int)(u - row * upr);
        const uint4 *src = in + (row * upr + (upr - 1 - ux)) * (32 * K);
#pragma unroll
        for (int j = 0; j < K; ++j) stage[wib][lane + 32 * j] = ld_stream(src + lane + 32 * j);
        __syncwarp();
        // lane takes the mirrored 4-pixel block (K vectors) and reverses its pixels
        uint4 v[K];
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = stage[wib][(31 - lane) * K + k];
        __syncwarp();
        const uint32_t *w = reinterpret_cast<const uint32_t *>(v);
        uint4 o[K];
        uint32_t *ow = reinterpret_cast<uint32_t *>(o);
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int c = 0; c < K; ++c) ow[p * K + c] = w[(3 - p) * K + c];
#pragma unroll
        for (int k = 0; k < K; ++k) stage[wib][lane * K + k] = o[k];
        __syncwarp();
        uint4 *dst = out + (row * upr + ux) * (32 * K);
#pragma unroll
        for (int j = 0; j < K; ++j) st_stream(dst + lane + 32 * j, stage[wib][lane + 32 * j]);
        __syncwarp();
    }
}

// Scalar path for ragged widths: one 32-bit word per thread, coalesced stores.
template <int K>
__global__ void __launch_bounds__(256)
fliplr_scalar_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int width, size_t nwords)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t row_words = (size_t)width * K;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += stride) {
        size_t row = i / row_words;
        int j = (int)(i - row * row_words);
        int x = j / K, c = j - x * K;
        out[i] = in[row * row_words + (size_t)(width - 1 - x) * K + c];
    }
}

// ------------------------------------------------------------ rotate, nearest
// The reference's rule (src/millipyde_image.cpp:114-139): inverse map about
// (W/2, H/2), int truncation toward zero, zero fill.  The expressions are kept
// in the reference's shape (and sin/cos evaluated on the device in fp64) so the
// compiler's FMA contraction and the resulting truncation match the reference's
// kernel bit for bit.
template <int K>
__global__ void __launch_bounds__(256)
rotate_nearest_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int width, int height,
                      double angle)
{
    int x = threadIdx.x + blockIdx.x * blockDim.x;
    int y = threadIdx.y + blockIdx.y * blockDim.y;

    int x_rot = ((double)x - ((double)width / 2)) * cos(angle) -
                ((double)y - ((double)height / 2)) * sin(angle) + ((double)width / 2);
    int y_rot = ((double)x - ((double)width / 2)) * sin(angle) +
                ((double)y - ((double)height / 2)) * cos(angle) + ((double)height / 2);

    if (x < width && y < height) {
        size_t o = ((size_t)y * width + x) * K;
        if (x_rot >= 0 && x_rot < width && y_rot >= 0 && y_rot < height) {
            size_t s = ((size_t)y_rot * width + x_rot) * K;
#pragma unroll
            for (int c = 0; c < K; ++c) out[o + c] = __ldg(in + s + c);
        } else {
#pragma unroll
            for (int c = 0; c < K; ++c) out[o + c] = 0u;
        }
    }
}

// ----------------------------------------------------------- rotate, bilinear
// skimage.transform.rotate defaults (order=1, mode='constant', cval=0, centre
// (W/2-0.5, H/2-0.5)).  Source coordinates in fp64 (fp32 would carry ~2.4e-4 px
// at x ~ 3840, 24x the tolerance), blend in the image's precision.  Each corner
// is the pixel if inside, else 0.  A warp covers a 32 x 1 run of output pixels
// and a block a 32 x 8 patch, so the four-corner gathers of neighbouring lanes
// fall in the same few 128-byte lines and are served by L1 after the first
// touch (the read-only path, __ldg).

template <typename T, int C>
__global__ void __launch_bounds__(256)
rotate_bilinear_kernel(const T *__restrict__ in, T *__restrict__ out, int width, int height,
                       RotateParams rp)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= width || y >= height) return;
    const double fx = (double)x - rp.cx, fy = (double)y - rp.cy;
    const double xs = rp.c * fx - rp.s * fy + rp.cx;
    const double ys = rp.s * fx + rp.c * fy + rp.cy;
    const double xf = floor(xs), yf = floor(ys);
    const int x0 = (int)xf, y0 = (int)yf;
    const int x1 = (int)ceil(xs), y1 = (int)ceil(ys);
    const T dx = (T)(xs - xf), dy = (T)(ys - yf);
    const bool in_x0 = x0 >= 0 && x0 < width, in_x1 = x1 >= 0 && x1 < width;
    const bool in_y0 = y0 >= 0 && y0 < height, in_y1 = y1 >= 0 && y1 < height;
    const T *r0 = in + (size_t)(in_y0 ? y0 : 0) * width * C;
    const T *r1 = in + (size_t)(in_y1 ? y1 : 0) * width * C;
    const size_t c0 = (size_t)(in_x0 ? x0 : 0) * C, c1 = (size_t)(in_x1 ? x1 : 0) * C;
    T *o = out + ((size_t)y * width + x) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        T p00 = (in_y0 && in_x0) ? __ldg(r0 + c0 + c) : (T)0;
        T p01 = (in_y0 && in_x1) ? __ldg(r0 + c1 + c) : (T)0;
        T p10 = (in_y1 && in_x0) ? __ldg(r1 + c0 + c) : (T)0;
        T p11 = (in_y1 && in_x1) ? __ldg(r1 + c1 + c) : (T)0;
        T top = ((T)1 - dx) * p00 + dx * p01;
        T bot = ((T)1 - dx) * p10 + dx * p11;
        o[c] = ((T)1 - dy) * top + dy * bot;
    }
}

}  // namespace mpk

namespace mpk {

// ------------------------------------------------------- fused gather segment
// One pass for a run of index ops, at most one bilinear rotate, and the pointwise
// ops around it (fp32 HWC).  For output pixel p:
//     q      = post(p)                       flips that come AFTER the rotate, as an index map
//     (ys,xs)= Rot(q)  (or q itself)         fp64 source coordinates in the rotate's space
//     corner -> pre(corner)                  flips that come BEFORE the rotate
//     v      = sum_corners wgt * [inside ? pw_pre(src[pre(corner)]) : 0]
//     out    = pw_post(v)
// which is exactly what running the ops one after another computes (pointwise ops
// commute with index maps; the rotate's zero fill is applied to the already
// pre-processed image, so an outside corner contributes 0, not pw_pre(0)).
// Images of a batch are addressed through pointer tables; blockIdx.z is the image.
template <int C>
__global__ void __launch_bounds__(256)
gather_f32_kernel(const __grid_constant__ GatherParams g)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= g.out_w || y >= g.out_h) return;
    const float *__restrict__ src = g.in_tab ? g.in_tab[blockIdx.z] : g.in;
    float *__restrict__ dst = g.out_tab ? g.out_tab[blockIdx.z] : g.out;

    const int qy = g.post.ay * y + g.post.by * x + g.post.cy;
    const int qx = g.post.ax * y + g.post.bx * x + g.post.cx;
    float acc[C];
    auto fetch = [&](int cy, int cx, int c) -> float {
        const int sy = g.pre.ay * cy + g.pre.by * cx + g.pre.cy;
        const int sx = g.pre.ax * cy + g.pre.bx * cx + g.pre.cx;
        return pw_apply<C>(g.pw_pre, __ldg(src + ((size_t)sy * g.src_w + sx) * C + c), c);
    };
    if (g.has_rotate) {
        const double fx = (double)qx - g.rp.cx, fy = (double)qy - g.rp.cy;
        const double xs = g.rp.c * fx - g.rp.s * fy + g.rp.cx;
        const double ys = g.rp.s * fx + g.rp.c * fy + g.rp.cy;
        const double xf = floor(xs), yf = floor(ys);
        const int x0 = (int)xf, y0 = (int)yf, x1 = (int)ceil(xs), y1 = (int)ceil(ys);
        const float dx = (float)(xs - xf), dy = (float)(ys - yf);
        const bool in_x0 = x0 >= 0 && x0 < g.rot_w, in_x1 = x1 >= 0 && x1 < g.rot_w;
        const bool in_y0 = y0 >= 0 && y0 < g.rot_h, in_y1 = y1 >= 0 && y1 < g.rot_h;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float p00 = (in_y0 && in_x0) ? fetch(y0, x0, c) : 0.f;
            const float p01 = (in_y0 && in_x1) ? fetch(y0, x1, c) : 0.f;
            const float p10 = (in_y1 && in_x0) ? fetch(y1, x0, c) : 0.f;
            const float p11 = (in_y1 && in_x1) ? fetch(y1, x1, c) : 0.f;
            const float top = (1.f - dx) * p00 + dx * p01;
            const float bot = (1.f - dx) * p10 + dx * p11;
            acc[c] = (1.f - dy) * top + dy * bot;
        }
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = fetch(qy, qx, c);
    }
    float *o = dst + ((size_t)y * g.out_w + x) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) o[c] = pw_apply<C>(g.pw_post, acc[c], c);
}

}  // namespace mpk
