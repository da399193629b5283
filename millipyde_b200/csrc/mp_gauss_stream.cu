// Launcher of the fp32 streaming Gaussian (kernels/gaussian_stream.cuh).
//
// Radius buckets: the kernel is fully unrolled over the taps, so it is
// instantiated for a few radii and a request uses the smallest bucket that
// holds its effective radius (weights beyond the radius are zero).  sigma = 2
// (effective radius 10) lands exactly on a bucket.
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "kernels/gaussian_stream.cuh"
#include "kernels/gaussian_stream_ws.cuh"
#include "kernels/gaussian_stream_mma.cuh"
#include "mp_image.h"
#include "mp_internal.h"
#include "mp_ops_internal.h"

using namespace mpk;

namespace mp {

static const int kBuckets[] = {3, 5, 7, 9, 10, 11, 13};

static int bucket_for(int radius)
{
    for (int b : kBuckets)
        if (radius <= b) return b;
    return 0;
}

bool gauss_stream_supported(int W, int C, int radius)
{
    // any width: rows that are not whole vectors (W * C % 4 != 0) go through the *_sets kernels, which
    // land the enclosing aligned span and shift it into place
    return (size_t)W * C >= 64 && bucket_for(radius) != 0;
}

// Which column pass a launch uses: the tensor-core one (gaussian_stream_mma.cuh) wherever it is
// instantiated (an output block spans at most 4 chunks: R <= 11), unless mpimg_set_gauss_column /
// MILLIPYDE_GAUSS_COLUMN=fma asks for the FMA-pipe kernel (A/B measurements and parity tests; the
// two agree to ~1e-6, not bit for bit).
static std::atomic<int> g_column{-1};
static std::atomic<int> g_range{-1};
static int value_range()
{
    int m = g_range.load(std::memory_order_relaxed);
    if (m < 0) {
        const char *e = getenv("MILLIPYDE_VALUE_RANGE");
        m = (e && strcmp(e, "any") == 0) ? MP_RANGE_ANY : MP_RANGE_UNIT;
        g_range.store(m);
    }
    return m;
}
static bool use_mma_column()
{
    // fp16 correction operands: only for data declared to lie in [0, 1] (include/mp_image.h)
    if (value_range() == MP_RANGE_ANY) return false;
    int m = g_column.load(std::memory_order_relaxed);
    if (m < 0) {
        const char *e = getenv("MILLIPYDE_GAUSS_COLUMN");
        m = (e && strcmp(e, "fma") == 0) ? MP_GAUSS_COLUMN_FMA : MP_GAUSS_COLUMN_MMA;
        g_column.store(m);
    }
    return m == MP_GAUSS_COLUMN_MMA;
}

// Row steps a work item of n_out output rows costs its CTA: the rows it filters (n_out + 2R halo
// rows) rounded the way the kernels round them (kernels/gaussian_stream_ws.cuh: ws_steps).
static long item_row_steps(int n_out, int R, bool mma)
{
    const int n_rows = n_out + 2 * R;
    if (!mma) return (long)((n_rows + 9) / 10) * 10;
    return (long)((((n_rows + 7 + 11) / 12) + 3) & ~3) * 12;
}

// Work items of one launch.  A column is (image, strip), `columns` of them, `height` rows each; the
// grid is one persistent CTA per SM and CTA b takes items b, b + sms, ...
//   main part: the first floor(columns / sms) * sms columns, one item each -- whole waves, no halo
//       rows re-filtered;
//   tail: the remaining columns (< sms) cut into c row chunks each.  Every chunk re-filters 2R halo
//       rows and pads its row count to the hand-off ring, but the tail's makespan is what the whole
//       grid waits for: c minimises waves(c) * steps(rows / c).  (256 4K RGB images on 148 SMs: 4608
//       columns = 31 waves + 20 columns; as 20 more whole columns they cost a 32nd wave with 128 SMs
//       idle, as 140 chunks of 309 rows a sixth of one.)
// A launch with fewer columns than SMs is all tail (a single 4K image: 18 columns x 8 chunks).
void gs_plan_items(long columns, int height, int R, int sms, bool mma, int *main_items, int *tail_cols,
                   int *chunk_rows, int *n_chunks)
{
    if (sms < 1) sms = 1;
    // MILLIPYDE_GAUSS_TAIL=0: every column is cut the same way (the round-1 schedule; A/B measurements)
    static const bool split_tail = [] { const char *e = getenv("MILLIPYDE_GAUSS_TAIL"); return !(e && *e == '0'); }();
    const long whole = split_tail ? columns / sms * sms : 0;
    const int tail = (int)(columns - whole);
    *main_items = (int)whole;
    *tail_cols = tail > 0 ? tail : 1;
    *chunk_rows = height;
    *n_chunks = 0;
    if (tail == 0) return;
    const int c_max = height / (4 * R) > 1 ? height / (4 * R) : 1;
    long best = 0;
    int chunks = 1;
    for (int c = 1; c <= c_max && c <= 64; ++c) {
        const int rows = (height + c - 1) / c;
        const int c_eff = (height + rows - 1) / rows;   // chunks that start inside the image
        const long waves = ((long)tail * c_eff + sms - 1) / sms;
        const long cost = waves * item_row_steps(rows, R, mma);
        if (best == 0 || cost * 100 < best * 98) {  // prefer fewer chunks unless the gain is real
            best = cost;
            chunks = c;
        }
    }
    *chunk_rows = (height + chunks - 1) / chunks;
    // rounding the chunk height up can leave the last chunks empty (770 rows in 64 chunks of 13 end
    // at chunk 59): count the chunks that start inside the image, every item must own at least one row
    *n_chunks = (height + *chunk_rows - 1) / *chunk_rows;
}

template <int C, int R>
static MPStatus launch_cr(int device, cudaStream_t s, GaussStreamParams &p, const GaussWeightSets *sets)
{
    static const int col_wait = [] { const char *e = getenv("MILLIPYDE_GAUSS_COLWAIT"); return e && *e == '1' ? 1 : 0; }();
    p.col_wait = col_wait;
    // MILLIPYDE_GAUSS_WAIT="a,b,c": sleep between probes of the three blocked waits (kernels/gaussian_stream.cuh:
    // wait_ns); a value with a leading 's' asks for a suspended try_wait with that hint (A/B measurements)
    static const struct WaitCfg {
        unsigned v[3];
        WaitCfg()
        {
            v[0] = v[1] = v[2] = kWsSleepNs;
            const char *e = getenv("MILLIPYDE_GAUSS_WAIT");
            for (int i = 0; e && *e && i < 3; ++i) {
                unsigned flag = 0;
                if (*e == 's') { flag = 0x40000000u; ++e; }
                char *end = nullptr;
                const unsigned long n = strtoul(e, &end, 10);
                if (end == e) break;
                v[i] = (unsigned)(n & 0x3fffffffu) | flag;
                e = *end == ',' ? end + 1 : end;
            }
        }
    } wait_cfg;
    for (int i = 0; i < 3; ++i) p.wait_ns[i] = wait_cfg.v[i];
    const int sms = sm_count(device) ? sm_count(device) : 148;
    constexpr bool kHasMma = MmGeom<C, R>::NCH <= 4;
    const bool mma = kHasMma && use_mma_column();
    const size_t smem = mma ? MmGeom<C, R>::SMEM : WsGeom<C, R>::SMEM;
    // per device and per instantiation; two workers racing here both set the same attributes
    static std::atomic<bool> configured[64] = {};
    if (device >= 0 && device < 64 && !configured[device].load(std::memory_order_acquire)) {
        MP_CUDA_TRY(cudaFuncSetAttribute(gauss_stream_ws_kernel<C, R>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WsGeom<C, R>::SMEM));
        MP_CUDA_TRY(cudaFuncSetAttribute(gauss_stream_ws_sets_kernel<C, R>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WsGeom<C, R>::SMEM));
        if constexpr (kHasMma) {
            MP_CUDA_TRY(cudaFuncSetAttribute(gauss_stream_mma_kernel<C, R>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MmGeom<C, R>::SMEM));
            MP_CUDA_TRY(cudaFuncSetAttribute(gauss_stream_mma_sets_kernel<C, R>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MmGeom<C, R>::SMEM));
        }
        configured[device].store(true, std::memory_order_release);
    }
    // One persistent CTA per SM, items dealt round-robin (gs_plan_items below)
    gs_plan_items((long)p.n_images * p.n_strips, p.height, R, sms, mma, &p.main_items, &p.tail_cols, &p.chunk_rows,
                  &p.n_chunks);
    const long items = (long)p.main_items + (long)p.tail_cols * p.n_chunks;
    const int grid = (int)(items < sms ? items : sms);
    if constexpr (kHasMma) {
        if (mma) {
            if (sets) gauss_stream_mma_sets_kernel<C, R><<<grid, kMmThreads, smem, s>>>(p, *sets);
            else gauss_stream_mma_kernel<C, R><<<grid, kMmThreads, smem, s>>>(p);
            count_launch();
            return MILLIPYDE_SUCCESS;
        }
    }
    if (sets) gauss_stream_ws_sets_kernel<C, R><<<grid, kWsThreads, smem, s>>>(p, *sets);
    else gauss_stream_ws_kernel<C, R><<<grid, kWsThreads, smem, s>>>(p);
    count_launch();
    return MILLIPYDE_SUCCESS;
}

template <int C>
static MPStatus launch_c(int device, cudaStream_t s, GaussStreamParams &p, int bucket,
                         const GaussWeightSets *sets = nullptr)
{
    switch (bucket) {
        case 3: return launch_cr<C, 3>(device, s, p, sets);
        case 5: return launch_cr<C, 5>(device, s, p, sets);
        case 7: return launch_cr<C, 7>(device, s, p, sets);
        case 9: return launch_cr<C, 9>(device, s, p, sets);
        case 10: return launch_cr<C, 10>(device, s, p, sets);
        case 11: return launch_cr<C, 11>(device, s, p, sets);
        case 13: return launch_cr<C, 13>(device, s, p, sets);
        default: return MP_ERROR_INVALID_ARGUMENT;
    }
}

MPStatus launch_gauss_stream_batch(int device, cudaStream_t s, int H, int W, int C, int n_images,
                                   const float *in, float *out, size_t image_stride,
                                   const float *const *in_tab, float *const *out_tab,
                                   const GaussParams<float> &gp)
{
    const int bucket = bucket_for(gp.radius);
    if (!bucket) return MP_ERROR_INVALID_ARGUMENT;
    GaussStreamParams p = {};
    p.in = in;
    p.out = out;
    p.in_tab = in_tab;
    p.out_tab = out_tab;
    p.image_stride = image_stride;
    p.n_images = n_images;
    p.height = H;
    p.row_elems = W * C;
    p.n_strips = (p.row_elems + kGsTW - 1) / kGsTW;
    p.radius = gp.radius;
    for (int d = 0; d < 16; ++d) {
        p.w[d] = d <= gp.radius ? gp.w[d] : 0.f;
        unsigned int bits;
        memcpy(&bits, &p.w[d], 4);
        p.ww[d] = ((unsigned long long)bits << 32) | bits;
    }
    if (p.row_elems & 3) {   // only the *_sets kernels shift misaligned rows into place: every set = these weights
        static thread_local GaussWeightSets same_sets;
        for (int i = 0; i < kGsMaxSets; ++i)
            for (int d = 0; d < 14; ++d) same_sets.ww[i][d] = p.ww[d];
        p.radius = bucket;
        if (C == 1) return launch_c<1>(device, s, p, bucket, &same_sets);
        if (C == 3) return launch_c<3>(device, s, p, bucket, &same_sets);
        if (C == 4) return launch_c<4>(device, s, p, bucket, &same_sets);
        return MP_ERROR_UNSUPPORTED_LAYOUT;
    }
    if (C == 1) return launch_c<1>(device, s, p, bucket);
    if (C == 3) return launch_c<3>(device, s, p, bucket);
    if (C == 4) return launch_c<4>(device, s, p, bucket);
    return MP_ERROR_UNSUPPORTED_LAYOUT;
}

int gauss_stream_bucket(int radius) { return bucket_for(radius); }

// n_images <= kGsMaxSets images of one radius bucket, image i filtered with gps[i].
MPStatus launch_gauss_stream_sets(int device, cudaStream_t s, int H, int W, int C, int n_images,
                                  const float *const *in_tab, float *const *out_tab,
                                  const GaussParams<float> *gps, int gps_stride, const PwProgram *pw_tab,
                                  int pw_stride)
{
    // per-image weights: at most kGsMaxSets images (image i uses set i % kGsMaxSets); one sigma for all
    // (gps_stride 0): every set is the same, any number of images
    if (n_images < 1 || (gps_stride != 0 && n_images > kGsMaxSets)) return MP_ERROR_INVALID_ARGUMENT;
    int bucket = 0;
    for (int i = 0; i < (gps_stride ? n_images : 1); ++i) {
        const int b = bucket_for(gps[(size_t)i * gps_stride].radius);
        if (!b) return MP_ERROR_INVALID_ARGUMENT;
        if (b > bucket) bucket = b;
    }
    GaussStreamParams p = {};
    p.in_tab = in_tab;
    p.out_tab = out_tab;
    p.n_images = n_images;
    p.height = H;
    p.row_elems = W * C;
    p.n_strips = (p.row_elems + kGsTW - 1) / kGsTW;
    p.radius = bucket;
    p.pw_tab = pw_tab;
    p.pw_stride = pw_stride;
    static thread_local GaussWeightSets sets;  // 7 KB: keep it off the worker's stack frames
    for (int i = 0; i < (gps_stride ? n_images : kGsMaxSets); ++i)
        for (int d = 0; d < 14; ++d) {
            const GaussParams<float> &gi = gps[(size_t)i * gps_stride];
            const float w = d <= gi.radius ? gi.w[d] : 0.f;
            unsigned int bits;
            memcpy(&bits, &w, 4);
            sets.ww[i][d] = ((unsigned long long)bits << 32) | bits;
        }
    if (C == 1) return launch_c<1>(device, s, p, bucket, &sets);
    if (C == 3) return launch_c<3>(device, s, p, bucket, &sets);
    if (C == 4) return launch_c<4>(device, s, p, bucket, &sets);
    return MP_ERROR_UNSUPPORTED_LAYOUT;
}

MPStatus launch_gauss_stream(int device, cudaStream_t s, const Img &d, const float *in, float *out,
                             const GaussParams<float> &gp)
{
    return launch_gauss_stream_batch(device, s, d.H, d.W, d.C, 1, in, out, 0, nullptr, nullptr, gp);
}

}  // namespace mp

extern "C" {
void mpimg_set_gauss_column(int mode)
{
    mp::g_column.store(mode == MP_GAUSS_COLUMN_FMA ? MP_GAUSS_COLUMN_FMA : MP_GAUSS_COLUMN_MMA);
}
int mpimg_get_gauss_column(void) { return mp::use_mma_column() ? MP_GAUSS_COLUMN_MMA : MP_GAUSS_COLUMN_FMA; }
void mpimg_set_value_range(int mode) { mp::g_range.store(mode == MP_RANGE_ANY ? MP_RANGE_ANY : MP_RANGE_UNIT); }
int mpimg_get_value_range(void) { return mp::value_range(); }
void mpimg_gauss_stream_plan(long columns, int height, int radius, int sms, int mma, int out[4])
{
    mp::gs_plan_items(columns, height, radius, sms, mma != 0, &out[0], &out[1], &out[2], &out[3]);
}
}
