// Separable Gaussian, fp32 HWC, streaming kernel -- the roofline path.
//
// One HBM read and one HBM write per sample, both passes in one launch, no
// intermediate image.  Work item = (image, column strip of TW floats, row chunk);
// a persistent grid of CTAs walks the items.  Inside an item the CTA marches down
// the rows in steps of Q rows:
//
//   1. TMA:   one elected thread issues 1-D bulk copies (cp.async.bulk, SASS
//             UBLKCP) of the next rows' [x0-HALO, x0+TW+HALO) segments into a
//             2-stage shared-memory ring; completion is an mbarrier transaction
//             count, so no thread spends registers or issue slots on the loads.
//   2. rows:  warp q filters row q of the step horizontally.  A lane owns PH = 20
//             consecutive floats; it reads its 20 + 2*HALO window as aligned
//             LDS.128 (lane pitch 80 B = 5 x 16 B, odd => conflict-free) and
//             scatters every loaded sample into the <= 2R+1 accumulators it
//             feeds (taps are C floats apart: channels stay interleaved).  All
//             indices are compile-time, the weights are constant-bank operands.
//   3. cols:  thread t owns float columns 2t, 2t+1 for the whole item and keeps
//             the 2R+1 partially accumulated output rows of each in registers.
//             Each new filtered row costs ONE LDS.64 per thread: it is scattered
//             into the live accumulators, the oldest one is complete and leaves
//             as a coalesced 8-byte store.  The accumulator that retires is the
//             one the next row opens, so the register file acts as the ring; the
//             rotation is resolved at compile time by a switch over row % (2R+1).
//
// Per output sample: 2 x (2R+1) FMAs, ~7.4 shared-memory words, 8 HBM bytes.
// With R = 11 that is 46 FMA per 8 bytes = 5.75 flop/B x 2: above the fp32-pipe
// ridge of this part (~72 TFLOP/s / 6.4 TB/s = 11 flop/B), so the kernel is
// bound by the FMA pipe, not by HBM; DESIGN.md carries the arithmetic.
//
// Zero padding (scipy mode="constant", cval=0): rows outside the image are never
// loaded (their filtered row is written as zeros), columns outside are zeroed in
// the ring once per item.
//
// Requirements: (W*C) % 4 == 0 (16-byte row pitch for the bulk copies) and the
// effective radius within the instantiated buckets; everything else takes
// kernels/gaussian_tile.cuh.
#pragma once
#include "common.cuh"

namespace mpk {

constexpr int kGsQ = 10;            // rows per step = warps per CTA
constexpr int kGsThreads = 32 * kGsQ;
constexpr int kGsPH = 2 * kGsQ;     // floats per lane in the row pass
constexpr int kGsTW = 32 * kGsPH;   // strip width in floats (640)

template <int C, int R>
struct GsGeom {
    static constexpr int HALO = (R * C + 3) / 4 * 4;          // halo rounded up to 16 bytes
    static constexpr int ROW = kGsTW + 2 * HALO;              // floats per staged row
    static constexpr int NA = 2 * R + 1;                      // live output rows per column
    static constexpr size_t IN_BYTES = 2ull * kGsQ * ROW * 4;  // 2-stage input ring
    static constexpr size_t H_BYTES = 2ull * kGsQ * kGsTW * 4; // 2-stage filtered-row ring
    static constexpr size_t SMEM = IN_BYTES + H_BYTES + 64;
};

struct GaussStreamParams {
    const float *in;             // single image / contiguous batch base (or null)
    float *out;
    const float *const *in_tab;  // per-image pointers (device memory) when the batch is scattered
    float *const *out_tab;
    size_t image_stride;         // floats between images of a contiguous batch
    int n_images;
    int height;
    int row_elems;               // W * C
    int n_strips;                // ceil(row_elems / TW)
    int chunk_rows;              // rows per work item
    int n_chunks;
    int radius;                  // actual radius (<= R); w[d] = 0 beyond it
    float w[16];
    unsigned long long ww[16];   // (w[d], w[d]) packed for fma.rn.f32x2
};

// ---- packed fp32x2 arithmetic (SASS FFMA2 / FMUL2): one issue slot, two FMAs ----
__device__ __forceinline__ uint64_t pack2(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b)
{
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// ---- row pass: scatter one aligned window into PH/2 packed accumulators -------
// Output pair m = outputs (2m, 2m+1).  A tap whose offset k*C is even reads the
// pair (x[2j], x[2j+1]) exactly as LDS.128 delivered it; an odd offset needs the
// straddling pair (x[2j-1], x[2j]), rebuilt with two moves and reused by every odd
// tap.  All indices are compile-time; dead combinations vanish.
template <int C, int R>
__device__ __forceinline__ void gs_row_pass(const float *__restrict__ win, uint64_t (&acc)[kGsPH / 2],
                                            const GaussStreamParams &p)
{
    constexpr int HALO = GsGeom<C, R>::HALO;
    constexpr int NV = (kGsPH + 2 * HALO) / 4;
#pragma unroll
    for (int i = 0; i < kGsPH / 2; ++i) acc[i] = 0ull;
    uint64_t prev = 0ull;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const ulonglong2 ld = *reinterpret_cast<const ulonglong2 *>(win + 4 * v);
        const uint64_t e[2] = {ld.x, ld.y};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int j = 2 * v + u;  // this pair holds window offsets 2j, 2j+1
            float a_lo, a_hi, b_lo, b_hi;
            unpack2(prev, a_lo, a_hi);
            unpack2(e[u], b_lo, b_hi);
            const uint64_t odd = pack2(a_hi, b_lo);  // window offsets 2j-1, 2j
#pragma unroll
            for (int k = -R; k <= R; ++k) {
                const int kc = k * C;
                if ((kc & 1) == 0) {
                    const int m2 = 2 * j - HALO - kc;
                    if (m2 >= 0 && m2 < kGsPH) acc[m2 / 2] = ffma2(p.ww[k < 0 ? -k : k], e[u], acc[m2 / 2]);
                } else {
                    const int m2 = 2 * j - 1 - HALO - kc;
                    if (m2 >= 0 && m2 < kGsPH) acc[m2 / 2] = ffma2(p.ww[k < 0 ? -k : k], odd, acc[m2 / 2]);
                }
            }
            prev = e[u];
        }
    }
}

// ---- column pass: one filtered row into the register ring --------------------
// a[s] holds the two columns of this thread packed.  PHASE = (row index) mod NA;
// the slot of output row (r + d) is (PHASE + d) mod NA.
template <int R, int PHASE>
__device__ __forceinline__ uint64_t gs_col_row(uint64_t (&a)[2 * R + 1], uint64_t v, const GaussStreamParams &p)
{
    constexpr int NA = 2 * R + 1;
#pragma unroll
    for (int d = -R; d < R; ++d) {
        const int s = (PHASE + d + NA) % NA;
        a[s] = ffma2(p.ww[d < 0 ? -d : d], v, a[s]);
    }
    // the output row that opens at this input row starts its sum here
    a[(PHASE + R) % NA] = fmul2(p.ww[R], v);
    return a[(PHASE - R + NA) % NA];  // output row r - R is complete
}

// Jump table over the NA rotations (one indirect branch per row instead of a
// compare chain).
template <int R>
__device__ __forceinline__ uint64_t gs_col_dispatch(int phase, uint64_t (&a)[2 * R + 1], uint64_t v,
                                                    const GaussStreamParams &p)
{
    constexpr int NA = 2 * R + 1;
#define MP_GS_CASE(P)                                         \
    case P:                                                   \
        if constexpr (P < NA) return gs_col_row<R, (P < NA ? P : 0)>(a, v, p); \
        break;
    switch (phase) {
        MP_GS_CASE(0) MP_GS_CASE(1) MP_GS_CASE(2) MP_GS_CASE(3) MP_GS_CASE(4) MP_GS_CASE(5) MP_GS_CASE(6)
        MP_GS_CASE(7) MP_GS_CASE(8) MP_GS_CASE(9) MP_GS_CASE(10) MP_GS_CASE(11) MP_GS_CASE(12) MP_GS_CASE(13)
        MP_GS_CASE(14) MP_GS_CASE(15) MP_GS_CASE(16) MP_GS_CASE(17) MP_GS_CASE(18) MP_GS_CASE(19)
        MP_GS_CASE(20) MP_GS_CASE(21) MP_GS_CASE(22) MP_GS_CASE(23) MP_GS_CASE(24) MP_GS_CASE(25)
        MP_GS_CASE(26)
        default: break;
    }
#undef MP_GS_CASE
    return 0ull;
}

template <int C, int R>
__global__ void __launch_bounds__(kGsThreads, 2)
gauss_stream_kernel(const __grid_constant__ GaussStreamParams p)
{
    using G = GsGeom<C, R>;
    constexpr int NA = G::NA;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_in = reinterpret_cast<float *>(smem_raw);                    // [2][Q][ROW]
    float *s_h = reinterpret_cast<float *>(smem_raw + G::IN_BYTES);        // [2][Q][TW]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + G::IN_BYTES + G::H_BYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&bars[0], kGsQ);
        mbar_init(&bars[1], kGsQ);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint32_t parity_bits = 0;  // bit s = parity the next wait on stage s expects
    const int items_per_image = p.n_strips * p.n_chunks;
    const long n_items = (long)p.n_images * items_per_image;

    for (long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int img = (int)(item / items_per_image);
        const int rem = (int)(item - (long)img * items_per_image);
        // strips vary fastest so CTAs that run together share halo columns in L2
        const int chunk = rem / p.n_strips, strip = rem - chunk * p.n_strips;
        const float *__restrict__ src = p.in_tab ? p.in_tab[img] : p.in + (size_t)img * p.image_stride;
        float *__restrict__ dst = p.out_tab ? p.out_tab[img] : p.out + (size_t)img * p.image_stride;

        const int x0 = strip * kGsTW;                 // first output column (floats)
        const int y0 = chunk * p.chunk_rows;
        const int y1 = min(p.height, y0 + p.chunk_rows);
        const int r_begin = y0 - R;                   // first filtered row needed
        const int n_rows = (y1 - y0) + 2 * R;
        const int n_steps = (n_rows + kGsQ - 1) / kGsQ;

        // columns of the staged row that exist in the image: [lo, hi) in ring coordinates
        const int gx_start = x0 - G::HALO;
        const int lo = gx_start < 0 ? -gx_start : 0;
        const int hi = min(G::ROW, p.row_elems - gx_start);
        const uint32_t row_bytes = (uint32_t)(hi - lo) * 4u;

        // all reads of the previous item are done; zero the never-copied border columns
        __syncthreads();
        if (lo > 0 || hi < G::ROW) {
            for (int i = tid; i < 2 * kGsQ * G::ROW; i += kGsThreads) {
                const int col = i % G::ROW;
                if (col < lo || col >= hi) s_in[i] = 0.f;
            }
            fence_proxy_async();
            __syncthreads();
        }

        // Stage fill, spread over the CTA: lane 0 of warp q posts the expected bytes of row q and
        // issues its one bulk copy (10 arrivals per phase), so no single warp carries the issue work.
        auto issue = [&](int step) {  // lane 0 of every warp
            const int stage = step & 1;
            const int r = r_begin + step * kGsQ + warp;
            const bool live = r >= 0 && r < p.height && r < r_begin + n_rows;
            mbar_expect_tx(&bars[stage], live ? row_bytes : 0u);
            if (live)
                bulk_g2s(s_in + ((size_t)stage * kGsQ + warp) * G::ROW + lo,
                         src + (size_t)r * p.row_elems + gx_start + lo, row_bytes, &bars[stage]);
        };
        if (lane == 0) {
            issue(0);
            if (n_steps > 1) issue(1);
        }

        uint64_t a[NA];
#pragma unroll
        for (int i = 0; i < NA; ++i) a[i] = 0ull;
        int phase = 0;  // (row - r_begin) mod NA

        for (int step = 0; step < n_steps; ++step) {
            const int stage = step & 1;
            mbar_wait(&bars[stage], (parity_bits >> stage) & 1u);
            parity_bits ^= 1u << stage;

            // ---- row pass: warp = row of the step, lane = 20-float segment
            {
                const int r = r_begin + step * kGsQ + warp;
                float *hrow = s_h + ((size_t)stage * kGsQ + warp) * kGsTW + lane * kGsPH;
                uint64_t acc[kGsPH / 2];
                if (r >= 0 && r < p.height && r < r_begin + n_rows) {
                    const float *win = s_in + ((size_t)stage * kGsQ + warp) * G::ROW + lane * kGsPH;
                    gs_row_pass<C, R>(win, acc, p);
                } else {
#pragma unroll
                    for (int i = 0; i < kGsPH / 2; ++i) acc[i] = 0ull;
                }
#pragma unroll
                for (int v = 0; v < kGsPH / 4; ++v)
                    *reinterpret_cast<ulonglong2 *>(hrow + 4 * v) = make_ulonglong2(acc[2 * v], acc[2 * v + 1]);
            }
            __syncthreads();  // filtered rows visible; this stage's input rows are free
            if (lane == 0 && step + 2 < n_steps) issue(step + 2);

            // ---- column pass: thread = float columns 2*tid, 2*tid+1
            const int gx = x0 + 2 * tid;
            const bool col_ok = gx < p.row_elems;
#pragma unroll 1
            for (int q = 0; q < kGsQ; ++q) {
                const int r = r_begin + step * kGsQ + q;
                const uint64_t v =
                    *reinterpret_cast<const uint64_t *>(s_h + ((size_t)stage * kGsQ + q) * kGsTW + 2 * tid);
                const uint64_t o = gs_col_dispatch<R>(phase, a, v, p);
                phase = (phase + 1 == NA) ? 0 : phase + 1;
                const int orow = r - R;
                if (col_ok && orow >= y0 && orow < y1) {
                    float o_lo, o_hi;
                    unpack2(o, o_lo, o_hi);
                    __stcs(reinterpret_cast<float2 *>(dst + (size_t)orow * p.row_elems + gx), make_float2(o_lo, o_hi));
                }
            }
        }
    }
}

}  // namespace mpk
