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
// TH = rows per tile: 64 for three- and four-word pixels (24-33 KB in flight per CTA).
template <int K, int TH = 32>
__global__ void __launch_bounds__(256)
transpose_tma_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int width, int height,
                     const uint32_t *const *__restrict__ in_tab = nullptr,
                     uint32_t *const *__restrict__ out_tab = nullptr)
{
    constexpr int P = 32 * K + 4;
    __shared__ __align__(128) uint32_t tile[TH * P];
    __shared__ __align__(8) uint64_t bar;
    if (in_tab) {
        in = in_tab[blockIdx.z];
        out = out_tab[blockIdx.z];
    }
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * TH;
    const int tw = min(32, width - x0), th = min(TH, height - y0);
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    {
        // Lane 0 of warp w issues rows w, w + 8, ... from uniform registers: the eight warps issue side by
        // side (one warp issuing all the copies of a tile serially took longer than the copies), and
        // per-lane addresses would cost an election loop per copy.  Thread 0 posts the byte count; a
        // copy that completes before that only drives the count negative meanwhile.
        const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);
        if ((tid & 31) == 0) {
            const uint32_t row_bytes = (uint32_t)tw * K * 4u;
            if (w == 0) mbar_expect_tx(&bar, row_bytes * (uint32_t)th);
            const uint32_t bar32 = smem_addr(&bar);
            uint32_t sdst = smem_addr(tile) + (uint32_t)(w * P) * 4u;
            const uint32_t *gsrc = in + ((size_t)(y0 + w) * width + x0) * K;
            const size_t gstep = (size_t)8 * width * K;
            for (int y = w; y < th; y += 8, gsrc += gstep, sdst += 8u * P * 4u) bulk_g2s_raw(sdst, gsrc, row_bytes, bar32);
            if (w == 0) mbar_wait(&bar, 0);   // one poller; the other warps park at the barrier below
        }
    }
    __syncthreads();
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

// One- and two-word pixels (fp32 greyscale, packed RGBA8; fp64 greyscale): 64-word x 64-row tiles
// and a skewed shared-memory layout.  With 4-byte pixels the 32 x 32 tile above is 4 KB -- too little
// in flight per CTA to cover the TMA round trip (50 % of the HBM peak) -- and its gather conflicts
// 4-way: a warp reads rows 4q + e for q = 0..7 of columns r..r+3, and any 16-byte-aligned row pitch
// puts rows 4 apart 16 banks apart at best.  Here row y lands at y * 96 + 4 * ((y >> S) & 7) words
// (S = 2 for one-word pixels, 1 for two-word pixels): the rows a warp gathers from start 4 banks
// apart, its columns fill the gaps, every LDS.32 (K = 1) / LDS.64 (K = 2) hits distinct banks, and
// 16 KB per CTA are in flight.  Stores are unchanged: 8 consecutive lanes write 128 contiguous bytes
// of one output row.  Same contract as transpose_tma_kernel<K> ((W*K) % 4 == 0 and (H*K) % 4 == 0).
constexpr int kT64Rows = 64;
constexpr int kT64Pitch = 96;
template <int K>
__global__ void __launch_bounds__(256)
transpose_tma64_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int width, int height,
                       const uint32_t *const *__restrict__ in_tab = nullptr,
                       uint32_t *const *__restrict__ out_tab = nullptr)
{
    static_assert(K == 1 || K == 2, "one- and two-word pixels");
    constexpr int TX = 64 / K;            // tile width in pixels (64 words)
    constexpr int S = K == 1 ? 2 : 1;     // rows per output vector = 4 / K = 1 << S
    __shared__ __align__(128) uint32_t tile[kT64Rows * kT64Pitch];
    __shared__ __align__(8) uint64_t bar;
    if (in_tab) {
        in = in_tab[blockIdx.z];
        out = out_tab[blockIdx.z];
    }
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * kT64Rows;
    const int tw = min(TX, width - x0), th = min(kT64Rows, height - y0);
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    {
        // lane 0 of warp w issues rows w, w + 8, ... (see transpose_tma_kernel)
        const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);
        if ((tid & 31) == 0) {
            const uint32_t row_bytes = (uint32_t)tw * K * 4u;
            if (w == 0) mbar_expect_tx(&bar, row_bytes * (uint32_t)th);
            const uint32_t bar32 = smem_addr(&bar), tile32 = smem_addr(tile);
            const uint32_t *gsrc = in + ((size_t)(y0 + w) * width + x0) * K;
            const size_t gstep = (size_t)8 * width * K;
            for (int y = w; y < th; y += 8, gsrc += gstep)
                bulk_g2s_raw(tile32 + (uint32_t)(y * kT64Pitch + 4 * ((y >> S) & 7)) * 4u, gsrc, row_bytes, bar32);
            if (w == 0) mbar_wait(&bar, 0);   // one poller; the other warps park at the barrier below
        }
    }
    __syncthreads();
    // output row (x0 + r) holds pixels y0 .. y0+th-1 = th * K / 4 vectors; a warp handles 8 vectors of 4 rows
    const int vpr = th * K / 4;
    const int q_groups = (vpr + 7) / 8;
    for (int i = tid; i < 8 * tw * q_groups; i += 256) {
        const int q_lo = i & 7, t2 = i >> 3;
        const int q_hi = t2 / tw, r = t2 - q_hi * tw;
        const int q = 8 * q_hi + q_lo;
        if (q >= vpr) continue;
        // vector q = input rows (q << S) .. ((q + 1) << S) - 1, all with skew 4 * (q & 7)
        const uint32_t *src = tile + (q << S) * kT64Pitch + 4 * q_lo + r * K;
        uint4 v;
        if (K == 1) {
            v = make_uint4(src[0], src[kT64Pitch], src[2 * kT64Pitch], src[3 * kT64Pitch]);
        } else {
            const uint2 a = *reinterpret_cast<const uint2 *>(src);
            const uint2 b = *reinterpret_cast<const uint2 *>(src + kT64Pitch);
            v = make_uint4(a.x, a.y, b.x, b.y);
        }
        st_stream(reinterpret_cast<uint4 *>(out + ((size_t)(x0 + r) * height + y0) * K) + q, v);
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

    // cos / sin of the launch's one angle: evaluated by the device's own fp64 routines (the values the
    // reference's kernel sees), but once per block instead of per pixel -- two ~50-instruction fp64
    // routines per thread were most of this kernel's time
    __shared__ double s_cs[2];
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        s_cs[0] = cos(angle);
        s_cs[1] = sin(angle);
    }
    __syncthreads();
    const double ca = s_cs[0], sa = s_cs[1];

    int x_rot = ((double)x - ((double)width / 2)) * ca -
                ((double)y - ((double)height / 2)) * sa + ((double)width / 2);
    int y_rot = ((double)x - ((double)width / 2)) * sa +
                ((double)y - ((double)height / 2)) * ca + ((double)height / 2);

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

// ------------------------------------------- rotate through a staged box (reference layouts)
// The same two rules -- the reference's nearest rule (RGBA8 words; fp64 greyscale under reference
// semantics as two words) and the oracle's bilinear rule on fp64 greyscale -- with the source
// footprint of a 32 x 64 (one-word pixels) or 32 x 32 (two-word pixels) output tile staged in shared
// memory by 1-D TMA bulk copies, one per box row, issued by lane 0 of all eight warps side by side:
// the direct kernels above walk a rotated line across ~18 cache lines per warp load.  (A first box
// version staged with per-thread loads and one issuing warp and was slower than the direct form.)
// Coordinates are evaluated per pixel by exactly the expressions of the direct kernels, so the
// results are the same bits; the box only changes where the sample is read from, and a coordinate
// the box does not hold -- it cannot happen with the margins below, but exactness must not depend
// on that -- is read from global memory.  Needs 16-byte-aligned rows ((W * K) % 4 == 0).
template <int K>
struct RotBoxGeom {
    static constexpr int TH = K == 1 ? 64 : 32;
    static constexpr int BOX = K == 1 ? 78 : 52;   // tile diagonal (70.2 / 43.9) + 2 + 2 margin + rounding
    static constexpr int PITCH = (BOX * K + 3 + 3) / 4 * 4 + 4;   // words; + 4: rows spread over the banks
    static constexpr size_t SMEM = (size_t)BOX * PITCH * 4;
};

template <int K, bool BILINEAR>
__global__ void __launch_bounds__(256)
rotate_box_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int width, int height, double angle,
                  RotateParams rp)
{
    using G = RotBoxGeom<K>;
    static_assert(!BILINEAR || K == 2, "bilinear: fp64 greyscale");
    constexpr int PITCH = G::PITCH, BOX = G::BOX, NPX = G::TH / 8;
    extern __shared__ __align__(16) uint32_t rbox[];
    __shared__ uint64_t s_bar;
    __shared__ double s_cs[2];
    __shared__ int s_geo[4];
    const int lane = threadIdx.x & 31, w = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (!BILINEAR) {   // the device's own fp64 routines, as the reference's kernel evaluates them
            s_cs[0] = cos(angle);
            s_cs[1] = sin(angle);
        }
    }
    __syncthreads();
    const double ca = BILINEAR ? rp.c : s_cs[0], sa = BILINEAR ? rp.s : s_cs[1];
    const int ox0 = blockIdx.x * 32, oy0 = blockIdx.y * G::TH;
    if (w == 0) {
        // the map is affine: the footprint's extremes are at the tile's corners
        const double cx = BILINEAR ? rp.cx : (double)width / 2, cy = BILINEAR ? rp.cy : (double)height / 2;
        const int ox1 = min(ox0 + 32, width) - 1, oy1 = min(oy0 + G::TH, height) - 1;
        double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double fx = (double)((q & 1) ? ox1 : ox0) - cx, fy = (double)((q & 2) ? oy1 : oy0) - cy;
            const double xr = fx * ca - fy * sa + cx, yr = fx * sa + fy * ca + cy;
            xmin = fmin(xmin, xr);
            xmax = fmax(xmax, xr);
            ymin = fmin(ymin, yr);
            ymax = fmax(ymax, yr);
        }
        const int bx0 = __double2int_rd(xmin) - 2, by0 = __double2int_rd(ymin) - 2;
        if (lane == 0) {
            s_geo[0] = bx0;
            s_geo[1] = by0;
            s_geo[2] = min(__double2int_ru(xmax) + 3 - bx0, BOX);
            s_geo[3] = min(__double2int_ru(ymax) + 3 - by0, BOX);
        }
    }
    __syncthreads();
    const int bx0 = s_geo[0], by0 = s_geo[1], bw = s_geo[2], bh = s_geo[3];
    // staged words of a row: [0, want); the part [c_lo, c_hi) of every box row and the box rows
    // [r_lo, r_hi) exist in the image (see gather_f32_kernel)
    const int row_len = width * K;
    const int shift = (bx0 * K) & 3;
    const int want = (shift + bw * K + 3) & ~3;
    const int col0 = bx0 * K - shift;
    const int c_lo = col0 < 0 ? -col0 : 0;
    const int c_hi = want < row_len - col0 ? want : row_len - col0;
    const int r_lo = by0 < 0 ? -by0 : 0;
    const int r_hi = bh < height - by0 ? bh : height - by0;
    const bool any = c_hi > c_lo && r_hi > r_lo;
    if (any && lane == 0) {
        const uint32_t row_bytes = (uint32_t)(c_hi - c_lo) * 4u;
        if (w == 0) mbar_expect_tx(&s_bar, row_bytes * (uint32_t)(r_hi - r_lo));
        const uint32_t bar32 = smem_addr(&s_bar);
        uint32_t sdst = smem_addr(rbox) + (uint32_t)((r_lo + w) * PITCH + c_lo) * 4u;
        const uint32_t *gsrc = in + ((long)(by0 + r_lo + w) * row_len + col0 + c_lo);
        for (int r = r_lo + w; r < r_hi; r += 8, gsrc += 8l * row_len, sdst += 8u * PITCH * 4u)
            bulk_g2s_raw(sdst, gsrc, row_bytes, bar32);
    }
    if (any && (c_lo > 0 || c_hi < want || r_lo > 0 || r_hi < bh)) {
        const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
        for (int r = w; r < bh; r += 8) {
            uint4 *brow = reinterpret_cast<uint4 *>(rbox + r * PITCH);
            if (r < r_lo || r >= r_hi) {
                for (int v = lane; v < (want >> 2); v += 32) brow[v] = zero4;
            } else {
                for (int v = lane; v < (c_lo >> 2); v += 32) brow[v] = zero4;
                for (int v = (c_hi >> 2) + lane; v < (want >> 2); v += 32) brow[v] = zero4;
            }
        }
    }
    if (any) {
        if (w == 0) mbar_wait_suspend(&s_bar, 0, 2000);
        __syncthreads();
    }
    const int x = ox0 + lane;
    if (x >= width) return;
    const uint32_t *origin = rbox + shift - by0 * PITCH - bx0 * K;   // box address of source pixel (0, 0)
    const int bx1 = bx0 + bw, by1 = by0 + bh;
#pragma unroll
    for (int k = 0; k < NPX; ++k) {
        const int y = oy0 + w + 8 * k;
        if (y >= height) break;
        uint32_t *o = out + ((size_t)y * width + x) * K;
        if (!BILINEAR) {
            // src/millipyde_image.cpp:114-139, the expressions of rotate_nearest_kernel
            int x_rot = ((double)x - ((double)width / 2)) * ca -
                        ((double)y - ((double)height / 2)) * sa + ((double)width / 2);
            int y_rot = ((double)x - ((double)width / 2)) * sa +
                        ((double)y - ((double)height / 2)) * ca + ((double)height / 2);
            uint32_t v[K];
#pragma unroll
            for (int c = 0; c < K; ++c) v[c] = 0u;
            if (x_rot >= 0 && x_rot < width && y_rot >= 0 && y_rot < height) {
                if (any && x_rot >= bx0 && x_rot < bx1 && y_rot >= by0 && y_rot < by1) {
                    const uint32_t *p = origin + y_rot * PITCH + x_rot * K;
#pragma unroll
                    for (int c = 0; c < K; ++c) v[c] = p[c];
                } else {
                    const uint32_t *p = in + ((size_t)y_rot * width + x_rot) * K;
#pragma unroll
                    for (int c = 0; c < K; ++c) v[c] = __ldg(p + c);
                }
            }
            if (K == 2) *reinterpret_cast<uint2 *>(o) = make_uint2(v[0], v[K - 1]);
            else o[0] = v[0];
        } else {
            // the expressions of rotate_bilinear_kernel<double, 1>
            const double fx = (double)x - rp.cx, fy = (double)y - rp.cy;
            const double xs = rp.c * fx - rp.s * fy + rp.cx;
            const double ys = rp.s * fx + rp.c * fy + rp.cy;
            const int x0 = __double2int_rd(xs), y0 = __double2int_rd(ys);
            const double dx = xs - (double)x0, dy = ys - (double)y0;
            const int x1 = x0 + (dx > 0.0), y1 = y0 + (dy > 0.0);
            double p00, p01, p10, p11;
            if (any && x0 >= bx0 && x1 < bx1 && y0 >= by0 && y1 < by1) {
                // outside-image corners are staged as zeros: the rule's cval
                const double *b = reinterpret_cast<const double *>(origin);
                constexpr int DP = PITCH / 2;
                p00 = b[y0 * DP + x0];
                p01 = b[y0 * DP + x1];
                p10 = b[y1 * DP + x0];
                p11 = b[y1 * DP + x1];
            } else {
                const double *src = reinterpret_cast<const double *>(in);
                const bool in_x0 = x0 >= 0 && x0 < width, in_x1 = x1 >= 0 && x1 < width;
                const bool in_y0 = y0 >= 0 && y0 < height, in_y1 = y1 >= 0 && y1 < height;
                p00 = (in_y0 && in_x0) ? __ldg(src + (size_t)y0 * width + x0) : 0.0;
                p01 = (in_y0 && in_x1) ? __ldg(src + (size_t)y0 * width + x1) : 0.0;
                p10 = (in_y1 && in_x0) ? __ldg(src + (size_t)y1 * width + x0) : 0.0;
                p11 = (in_y1 && in_x1) ? __ldg(src + (size_t)y1 * width + x1) : 0.0;
            }
            const double top = (1.0 - dx) * p00 + dx * p01;
            const double bot = (1.0 - dx) * p10 + dx * p11;
            *reinterpret_cast<double *>(o) = (1.0 - dy) * top + dy * bot;
        }
    }
}

// ----------------------------------------------------------- rotate, bilinear
// skimage.transform.rotate defaults (order=1, mode='constant', cval=0, centre
// (W/2-0.5, H/2-0.5)).  Source coordinates in fp64 (fp32 would carry ~2.4e-4 px
// at x ~ 3840, 24x the tolerance), blend in the image's precision.  Each corner
// is the pixel if inside, else 0.  This direct form (a warp covers a 32 x 1 run of
// output pixels, corners through the read-only path) serves the fp64 greyscale
// layout; fp32 images go through the tile-staged gather_f32_kernel below, whose
// loads are coalesced at every angle.
template <typename T, int C>
__global__ void __launch_bounds__(256)
rotate_bilinear_kernel(const T *__restrict__ in, T *__restrict__ out, int width, int height,
                       RotateParams rp)
{
    // a warp covers a 32 x 1 run of output pixels, the block 32 x 8 (measured: 8 x 4 warp patches
    // are slower here -- the stores dominate and want the long runs)
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= width || y >= height) return;
    const double fx = (double)x - rp.cx, fy = (double)y - rp.cy;
    const double xs = rp.c * fx - rp.s * fy + rp.cx;
    const double ys = rp.s * fx + rp.c * fy + rp.cy;
    // floor by cvt.rmi; ceil = floor + (frac > 0); both corner pairs share the fraction
    const int x0 = __double2int_rd(xs), y0 = __double2int_rd(ys);
    const T dx = (T)(xs - (double)x0), dy = (T)(ys - (double)y0);
    const int x1 = x0 + (dx > (T)0), y1 = y0 + (dy > (T)0);
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
// Tile form: a CTA produces a 32 x 32 output tile.  It first stages the tile's source footprint --
// the bounding box of the rotated tile, at most 50 x 50 pixels -- in shared memory: every row of
// the box is one contiguous global segment, so the loads are coalesced no matter the angle (a
// 32 x 1 run of output pixels walks a rotated line across ~18 cache lines per load instruction,
// which made the direct gather L1-wavefront-bound).  Outside-image samples are staged as zeros,
// which IS the rotate's cval = 0 rule.  The four corner reads per output sample hit shared memory.
//
// The kernel is issue-bound, so instructions are what is optimised:
//  * staging is done by the TMA unit: one 1-D bulk copy per box row, issued by the lanes of warp 0
//    (a lane owns rows lane and lane + 32), all completing on one mbarrier.  A row's copy starts at
//    the 16-byte boundary at or below its first float and ends at the one at or above its last;
//    image rows are whole vectors (checked: otherwise the scalar path stages), so neither leaves
//    the row and the alignment shift (0..3 floats) is the same for every row of the box.  Only
//    tiles whose box leaves the image zero anything, and only outside the copies' destinations.
//    (Maps with a flip or pointwise ops BEFORE the rotate use the scalar staging loop.)
//  * source coordinates: one fp64 affine evaluation per thread, then fp64 adds down the column;
//    floor is cvt.rmi, the blend runs in fp32.  The second corner is always the next pixel / next
//    row (weight exactly 0 when the coordinate is an integer; the box reaches ceil(max) + 1, so it
//    is staged), which makes the four corners one base address plus immediates.
//  * the thread's four pixels are blended first and the post-rotate pointwise program runs once
//    over all of them (its op decode is per program, not per pixel).
constexpr int kGatherTile = 32;      // tile width in output pixels (one per lane)
constexpr int kGatherBox = 50;       // box edge of a 32 x 32 tile: >= 32 * sqrt(2) + 4
constexpr int kGatherTileTall = 64;  // single-channel images: 32 x 64 tiles (a 32 x 32 x 4-byte tile is too little
constexpr int kGatherBoxTall = 76;   // work per CTA to amortise the set-up); box >= sqrt(32^2 + 64^2) + 4

template <int C, int TH = kGatherTile>
struct GatherGeom {
    static constexpr int BOX = TH == kGatherTile ? kGatherBox : kGatherBoxTall;
    // floats per staged row: box width * C plus up to 3 floats of alignment shift, rounded to a
    // vector, plus a skew that keeps PITCH % 32 in {4, 20} so rows spread over the banks
    static constexpr int ROW_MAX = (BOX * C + 3 + 3) / 4 * 4;
    static constexpr int PITCH = TH == kGatherTile ? (C == 1 ? 68 : (C == 3 ? 164 : 212)) : (C == 1 ? 84 : ROW_MAX + 4);
    static_assert(PITCH >= ROW_MAX && PITCH % 4 == 0, "pitch");
    static_assert(TH == kGatherTile || TH == kGatherTileTall, "tile heights");
    // The launcher may pick any pitch in [PITCH, PITCH_MAX] (a multiple of 4): which bank a corner read
    // of a rotated line hits is 3 * ix + pitch * iy, so the pitch with the fewest conflicts depends on
    // the angle (mp_image_ops.cu: gather_pitch_for)
    static constexpr int PITCH_MAX = PITCH + 16;
    static constexpr size_t SMEM = (size_t)BOX * PITCH_MAX * sizeof(float);
};

template <int C, bool TAB = false, int TH = kGatherTile>
__global__ void __launch_bounds__(256, 6)
gather_f32_kernel(const __grid_constant__ GatherParams g)
{
    const int PITCH = g.pitch ? g.pitch : GatherGeom<C, TH>::PITCH;
    constexpr int BOX = GatherGeom<C, TH>::BOX;
    constexpr int NPX = TH / 8;  // output pixels per thread
    extern __shared__ __align__(16) float box[];  // [bh][PITCH]
    // TAB: the angle and the pointwise programs are the image's own (GatherVar record in device
    // memory, copied to shared memory once per block); everything else is common to the launch
    __shared__ GatherVar s_var;
    __shared__ uint64_t s_bar;
    __shared__ int s_geo[4];  // the tile's source box (bx0, by0, bw, bh): computed by warp 0 for the block
    if (TAB) {
        for (int i = threadIdx.x; i < (int)(sizeof(GatherVar) / 4); i += blockDim.x)
            reinterpret_cast<int *>(&s_var)[i] = reinterpret_cast<const int *>(g.var_tab + blockIdx.z)[i];
    }
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const RotateParams &rp = TAB ? s_var.rp : g.rp;
    const PwProgram &pw_pre = TAB ? s_var.pw_pre : g.pw_pre;
    const PwProgram &pw_post = TAB ? s_var.pw_post : g.pw_post;
    const float *__restrict__ src = g.in_tab ? g.in_tab[blockIdx.z] : g.in;
    float *__restrict__ dst = g.out_tab ? g.out_tab[blockIdx.z] : g.out;
    const int ox0 = blockIdx.x * kGatherTile, oy0 = blockIdx.y * TH;
    // the warp index through a shuffle: the compiler then knows it is warp-uniform
    const int lane = threadIdx.x & 31, w = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);

    const bool identity_pre = g.pre.ay == 1 && g.pre.by == 0 && g.pre.cy == 0 && g.pre.ax == 0 &&
                              g.pre.bx == 1 && g.pre.cx == 0 && pw_pre.n == 0;
    const bool rows_aligned = ((g.src_w * C) & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    const bool tma = identity_pre && rows_aligned;
    const int row_len = g.src_w * C;

    // What follows from the box, in integers (every warp evaluates this once it knows the box):
    //   shift      floats between a staged row's start and box column 0 (same for every row)
    //   want       staged floats of a row: [0, want)
    //   col0       image-row float held by box-row float 0
    //   [c_lo, c_hi)  the part of every box row that exists in the image (multiples of 4)
    //   [r_lo, r_hi)  the box rows that lie inside the image
    int bx0, by0, bw, bh, shift = 0, want = 0, col0 = 0, c_lo = 0, c_hi = 0, r_lo = 0, r_hi = 0;
    bool any = true;
    auto derive = [&]() {
        if (!tma) return;
        shift = (bx0 * C) & 3;  // two's complement: right for bx0 < 0 too; row_len % 4 == 0
        want = (shift + bw * C + 3) & ~3;
        col0 = bx0 * C - shift;
        c_lo = col0 < 0 ? -col0 : 0;
        c_hi = want < row_len - col0 ? want : row_len - col0;
        r_lo = by0 < 0 ? -by0 : 0;
        r_hi = bh < g.rot_h - by0 ? bh : g.rot_h - by0;
        any = c_hi > c_lo && r_hi > r_lo;
    };

    if (w == 0) {
        // Footprint of the tile in the rotate's space (after the post map): both maps are affine, so
        // the image of the tile is a parallelogram around the image of its centre, and its bounding box
        // has half-extents |M| * (half-extents of the tile) -- a dozen fp64 operations instead of
        // mapping the four corners and reducing them.  One warp does it for the block (the kernel is
        // issue-bound: eight warps repeating it were an eighth of its instructions).
        const int ox1 = min(ox0 + kGatherTile, g.out_w) - 1, oy1 = min(oy0 + TH, g.out_h) - 1;
        const double hx = 0.5 * (ox1 - ox0), hy = 0.5 * (oy1 - oy0);
        const double pcx = ox0 + hx, pcy = oy0 + hy;
        const double qcy = g.post.ay * pcy + g.post.by * pcx + g.post.cy;
        const double qcx = g.post.ax * pcy + g.post.bx * pcx + g.post.cx;
        const double qhy = abs(g.post.ay) * hy + abs(g.post.by) * hx;
        const double qhx = abs(g.post.ax) * hy + abs(g.post.bx) * hx;
        double xc = qcx, yc = qcy, ex = qhx, ey = qhy;
        if (g.has_rotate) {
            const double fx = qcx - rp.cx, fy = qcy - rp.cy;
            xc = rp.c * fx - rp.s * fy + rp.cx;
            yc = rp.s * fx + rp.c * fy + rp.cy;
            ex = fabs(rp.c) * qhx + fabs(rp.s) * qhy;
            ey = fabs(rp.s) * qhx + fabs(rp.c) * qhy;
        }
        const double xmin = xc - ex, xmax = xc + ex, ymin = yc - ey, ymax = yc + ey;
        bx0 = __double2int_rd(xmin) - 1;
        by0 = __double2int_rd(ymin) - 1;
        bw = min(__double2int_ru(xmax) + 2 - bx0, BOX);
        bh = min(__double2int_ru(ymax) + 2 - by0, BOX);
        if (lane == 0) {
            s_geo[0] = bx0;
            s_geo[1] = by0;
            s_geo[2] = bw;
            s_geo[3] = bh;
        }
    }
    __syncthreads();  // the box is known to every warp
    if (w != 0) {
        bx0 = s_geo[0];
        by0 = s_geo[1];
        bw = s_geo[2];
        bh = s_geo[3];
    }
    derive();
    if (tma && any && lane == 0) {
        // ---- TMA staging: one 1-D bulk copy per box row that lies in the image, all on one mbarrier.
        // Lane 0 of warp w issues rows r_lo + w, + 8, ... from uniform registers: eight warps issue
        // side by side (one warp issuing ~50 copies one after the other took longer than the copies
        // themselves), and per-lane addresses would cost an election loop per copy.  Warp 0 posts the
        // byte count; a copy that completes before that only drives the count negative meanwhile.
        const uint32_t row_bytes = (uint32_t)(c_hi - c_lo) * 4u;
        if (w == 0) mbar_expect_tx(&s_bar, row_bytes * (uint32_t)(r_hi - r_lo));
        const uint32_t bar32 = smem_addr(&s_bar);
        uint32_t sdst = smem_addr(box) + (uint32_t)((r_lo + w) * PITCH + c_lo) * 4u;
        const float *gsrc = src + ((long)(by0 + r_lo + w) * row_len + col0 + c_lo);
        const long gstep = 8l * row_len;
        const uint32_t sstep = 8u * (uint32_t)PITCH * 4u;
        for (int r = r_lo + w; r < r_hi; r += 8, gsrc += gstep, sdst += sstep) bulk_g2s_raw(sdst, gsrc, row_bytes, bar32);
    }
    // a tile whose box misses the image: every corner of every sample is the rotate's cval
    const bool all_zero = tma && !any;

    if (tma) {
        // tiles whose box leaves the image zero what the copies do not cover: whole rows outside
        // [r_lo, r_hi), and the ends [0, c_lo) and [c_hi, want) of the rows inside
        if (any && (c_lo > 0 || c_hi < want || r_lo > 0 || r_hi < bh)) {
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = w; r < bh; r += 8) {
                float4 *brow = reinterpret_cast<float4 *>(box + r * PITCH);
                if (r < r_lo || r >= r_hi) {
                    for (int v = lane; v < (want >> 2); v += 32) brow[v] = zero4;
                } else {
                    for (int v = lane; v < (c_lo >> 2); v += 32) brow[v] = zero4;
                    for (int v = (c_hi >> 2) + lane; v < (want >> 2); v += 32) brow[v] = zero4;
                }
            }
        }
    } else {
        // ---- scalar staging through the pre map (and the pre-rotate pointwise ops)
        constexpr int PER_LANE = (BOX * C + 31) / 32;
        const bool has_pre = pw_pre.n > 0;
        const int row_floats = bw * C;
        for (int r = w; r < bh; r += 8) {
            const int cy = by0 + r;
            const bool row_in = cy >= 0 && cy < g.rot_h;
            float *brow = box + r * PITCH;
            float v[PER_LANE];
#pragma unroll
            for (int u = 0; u < PER_LANE; ++u) {
                const int j = lane + 32 * u;
                const int cxp = j / C, c = j - cxp * C;
                const int cx = bx0 + cxp;
                v[u] = 0.f;
                if (j < row_floats && row_in && cx >= 0 && cx < g.rot_w) {
                    const int sy = g.pre.ay * cy + g.pre.by * cx + g.pre.cy;
                    const int sx = g.pre.ax * cy + g.pre.bx * cx + g.pre.cx;
                    v[u] = __ldg(src + ((size_t)sy * g.src_w + sx) * C + c);
                    if (has_pre) v[u] = pw_apply<C>(pw_pre, v[u], c);
                }
            }
#pragma unroll
            for (int u = 0; u < PER_LANE; ++u) {
                const int j = lane + 32 * u;
                if (j < row_floats) brow[j] = v[u];
            }
        }
    }

    if (!all_zero) {
        // only the issuing warp polls the mbarrier; the others park at the CTA barrier, which costs no
        // issue slots (all 8 warps polling was 24 % of the kernel's instructions: 244 M probes per launch)
        if (tma && w == 0) mbar_wait_suspend(&s_bar, 0, 2000);
        __syncthreads();  // the copies have landed (warp 0 saw them) and so has every warp's staging / zero fill
    }
    const int x = ox0 + lane;
    if (x >= g.out_w) return;
    // output pixel (x, y): q = post(x, y); stepping y by 8 moves q by 8 * (post.ay, post.ax).  (Set up
    // after the barrier: held across it, the four fp64 values cost the 40-register budget spills.)
    const int y_first = oy0 + w;
    const int qy = g.post.ay * y_first + g.post.by * x + g.post.cy;
    const int qx = g.post.ax * y_first + g.post.bx * x + g.post.cx;
    const int dqy = 8 * g.post.ay, dqx = 8 * g.post.ax;
    double xs = qx, ys = qy, dxs = dqx, dys = dqy;
    if (g.has_rotate) {
        const double fx = (double)qx - rp.cx, fy = (double)qy - rp.cy;
        xs = rp.c * fx - rp.s * fy + rp.cx;
        ys = rp.s * fx + rp.c * fy + rp.cy;
        dxs = rp.c * dqx - rp.s * dqy;
        dys = rp.s * dqx + rp.c * dqy;
    }
    // pixels of this thread that exist: rows y_first + 8 k below the bottom edge
    const int n_live = y_first < g.out_h ? min(NPX, (g.out_h - y_first + 7) >> 3) : 0;
    const float *origin = box + shift - by0 * PITCH - bx0 * C;  // box address of source pixel (0, 0)
    float acc[NPX * C];
#pragma unroll
    for (int i = 0; i < NPX * C; ++i) acc[i] = 0.f;
    if (all_zero) {
        // nothing staged, nothing to read
    } else if (g.has_rotate) {
#pragma unroll
        for (int k = 0; k < NPX; ++k) {
            if (k < n_live) {
                // recomputing from k keeps every row one rounding away from the direct formula
                const double xk = xs + k * dxs, yk = ys + k * dys;
                const int ix = __double2int_rd(xk), iy = __double2int_rd(yk);
                const float dx = (float)(xk - (double)ix), dy = (float)(yk - (double)iy);
                const float *c00 = origin + iy * PITCH + ix * C;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float top = (1.f - dx) * c00[c] + dx * c00[C + c];
                    const float bot = (1.f - dx) * c00[PITCH + c] + dx * c00[PITCH + C + c];
                    acc[k * C + c] = (1.f - dy) * top + dy * bot;
                }
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < NPX; ++k) {
            if (k < n_live) {
                const float *c00 = origin + (qy + k * dqy) * PITCH + (qx + k * dqx) * C;
#pragma unroll
                for (int c = 0; c < C; ++c) acc[k * C + c] = c00[c];
            }
        }
    }
    pw_apply_tile<C, NPX * C>(pw_post, acc, 0);
    float *o = dst + ((size_t)y_first * g.out_w + x) * C;
    const size_t ostep = (size_t)8 * g.out_w * C;
#pragma unroll
    for (int k = 0; k < NPX; ++k, o += ostep) {
        if (k < n_live) {
#pragma unroll
            for (int c = 0; c < C; ++c) o[c] = acc[k * C + c];
        }
    }
}

}  // namespace mpk
