// Shared pieces of the fp32 streaming Gaussian (kernels/gaussian_stream_ws.cuh): strip geometry,
// the launch parameter block and the packed fp32x2 arithmetic (SASS FFMA2 / FMUL2).
//
// Geometry: a work item is (image, column strip of kGsTW = 640 floats, row chunk).  A lane of the
// row pass owns kGsPH = 20 consecutive floats (lane pitch 80 B = 5 x 16 B: odd, so the LDS.128 of a
// warp never conflict); a thread of the column pass owns 2 float columns.  HALO is the tap reach
// R*C rounded up to 16 bytes so every staged row starts on a vector boundary.
#pragma once
#include "common.cuh"

namespace mpk {

constexpr int kGsPH = 20;           // floats per lane in the row pass
constexpr int kGsTW = 32 * kGsPH;   // strip width in floats (640)

template <int C, int R>
struct GsGeom {
    static constexpr int HALO = (R * C + 3) / 4 * 4;          // halo rounded up to 16 bytes
    static constexpr int ROW = kGsTW + 2 * HALO;              // floats per staged row
    // slot pitch: one more vector, where the landing of a row that is not 16-byte aligned in global
    // memory (W * C % 4 != 0) spills before it is shifted into place
    static constexpr int SLOT = ROW + 4;
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
    // Work items.  A column is (image, strip); columns are numbered image-major.  The first
    // main_items columns are one item each (all rows); each of the tail_cols columns behind them is
    // cut into n_chunks row chunks of chunk_rows rows, numbered chunk-major (gs_item below).  The
    // launcher makes the main part whole waves of the persistent grid and cuts the remainder so that
    // it fills one more, short, wave (mp_gauss_stream.cu: gs_plan_items).
    int main_items;
    int tail_cols;
    int chunk_rows;              // rows per tail item
    int n_chunks;
    int radius;                  // actual radius (<= R); w[d] = 0 beyond it
    // Pointwise work fused around the blur (device memory, null = none): image i applies the program
    // pw_tab[2 i] to every input sample before the horizontal filter, and pw_tab[2 i + 1] -- n = 0, or
    // ONE op read as out = min(max(v + a, b), c) -- to every output sample (pw_stride = 2); or all
    // images share pw_tab[0], pw_tab[1] (pw_stride = 0).  Only the *_sets kernels look at it.
    const PwProgram *pw_tab;
    int pw_stride;
    int col_wait;                // COLUMN role's hand-off wait: 0 = probe + nanosleep, 1 = suspended try_wait (A/B switch)
    // nanoseconds a blocked wait sleeps between probes (tensor-core flavour): [0] a ROW warp waiting for
    // its TMA row, [1] a ROW warp waiting for the COLUMN warps to hand a ring group back, [2] a COLUMN
    // warp waiting for a filtered group.  Bit 30 set: a suspended try_wait with that time hint instead.
    unsigned wait_ns[3];
    float w[16];
    unsigned long long ww[16];   // (w[d], w[d]) packed for fma.rn.f32x2
};

// Work item -> (image, strip, rows [y0, y1)).  Evaluated once per item by every warp of both roles
// (two integer divisions per ~2000 rows of work).
struct GsItem {
    int img, strip, y0, y1;
};
__device__ __forceinline__ long gs_item_count(const GaussStreamParams &p)
{
    return (long)p.main_items + (long)p.tail_cols * p.n_chunks;
}
__device__ __forceinline__ GsItem gs_item(const GaussStreamParams &p, long item)
{
    GsItem it;
    int col;
    if (item < p.main_items) {
        col = (int)item;
        it.y0 = 0;
        it.y1 = p.height;
    } else {
        const int t = (int)(item - p.main_items);
        const int chunk = t / p.tail_cols;
        col = p.main_items + (t - chunk * p.tail_cols);
        it.y0 = chunk * p.chunk_rows;
        it.y1 = min(p.height, it.y0 + p.chunk_rows);
    }
    it.img = col / p.n_strips;
    it.strip = col - it.img * p.n_strips;
    return it;
}

// Per-image weight sets of a batched launch whose images share the radius bucket but not sigma
// (Generator streams with random_gaussian).  The table travels in the kernel parameter block, i.e.
// the constant bank: the image index of a work item is warp-uniform, so the weights still reach the
// FFMA2s as uniform-register operands.  Image i uses set i % kGsMaxSets (the launcher sends at most
// kGsMaxSets images per launch when the sets differ).
constexpr int kGsMaxSets = 64;
struct GaussWeightSets {
    unsigned long long ww[kGsMaxSets][14];  // (w[d], w[d]) for d = 0..13
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

}  // namespace mpk
