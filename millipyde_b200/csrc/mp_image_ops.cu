// The eight image operators of the hot path: layout dispatch + launches.
//
// Replaces the extern "C" block and the static launch wrappers of
// src/millipyde_image.cpp:530-1179.  Every op: selects obj->mem_loc, takes its
// output from the stream-ordered pool, launches on obj->stream, returns the old
// buffer to the pool in stream order (the reference hipMalloc'ed and hipFree'd
// -- an implicit device sync -- around every kernel), rewrites the MPObjData
// header where the shape or dtype changes, and returns a real MPStatus.
#include <cmath>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <type_traits>
#include <vector>
#include <map>
#include <mutex>
#include <string>

#include "kernels/gaussian_global.cuh"
#include "kernels/gaussian_tile.cuh"
#include "kernels/geometry.cuh"
#include "kernels/pointwise.cuh"
#include "mp_image.h"
#include "mp_internal.h"
#include "mp_ops_internal.h"

using namespace mpk;

namespace {

std::atomic<int> g_semantics{-1};

int semantics()
{
    int s = g_semantics.load(std::memory_order_relaxed);
    if (s < 0) {
        const char *e = getenv("MILLIPYDE_SEMANTICS");
        s = (e && strcmp(e, "reference") == 0) ? MP_SEMANTICS_REFERENCE : MP_SEMANTICS_ORACLE;
        g_semantics.store(s);
    }
    return s;
}

MPStatus begin(MPObjData *obj, mp::Img *d, cudaStream_t *stream)
{
    if (!obj) return MP_ERROR_INVALID_ARGUMENT;
    MPStatus st = mp::ensure_initialized();
    if (st != MILLIPYDE_SUCCESS) return st;
    if (obj->device_data == NULL) return MP_ERROR_NULL_DATA;
    if (!mp::describe(obj, d)) return MP_ERROR_UNSUPPORTED_LAYOUT;
    MP_CUDA_TRY(cudaSetDevice(obj->mem_loc));
    *stream = mp::stream_of(obj);
    return MILLIPYDE_SUCCESS;
}

// Output buffer from the pool; on success the op launches into it and then calls
// finish() which retires the input buffer in stream order.
MPStatus fresh(MPObjData *obj, cudaStream_t s, size_t nbytes, void **out)
{
    *out = mp::pool_alloc(obj->mem_loc, s, nbytes);
    return *out ? MILLIPYDE_SUCCESS : MP_ERROR_DEVICE_ALLOC;
}

MPStatus finish(MPObjData *obj, cudaStream_t s, void *out, size_t nbytes)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        mp::record_cuda_error(e, "kernel launch", __FILE__, __LINE__);
        mp::pool_free(obj->mem_loc, s, out);
        return MP_ERROR_CUDA_RUNTIME;
    }
    mp::pool_free(obj->mem_loc, s, obj->device_data);
    obj->device_data = out;
    obj->nbytes = nbytes;
    return MILLIPYDE_SUCCESS;
}

}  // namespace

namespace mp {

bool describe(const MPObjData *o, Img *d)
{
    if (o->ndims != 2 && o->ndims != 3) return false;
    d->H = o->dims[0];
    d->W = o->dims[1];
    d->C = o->ndims == 2 ? 1 : o->dims[2];
    d->type = o->type;
    if (d->H <= 0 || d->W <= 0 || d->C <= 0) return false;
    d->npix = (size_t)d->H * d->W;
    switch (o->type) {
        case MP_NPY_UBYTE: d->fam = (d->C == 4) ? FAM_RGBA8 : FAM_U8_OTHER; d->esize = 1; break;
        case MP_NPY_DOUBLE: d->fam = (d->C == 1) ? FAM_F64 : FAM_F64_OTHER; d->esize = 8; break;
        case MP_NPY_FLOAT:
            if (d->C != 1 && d->C != 3 && d->C != 4) return false;
            d->fam = FAM_F32;
            d->esize = 4;
            break;
        default: return false;
    }
    return true;
}

// words (32-bit) per pixel for the index kernels, 0 if the pixel is not word-sized
int words_per_pixel(const Img &d)
{
    size_t bytes = (size_t)d.C * d.esize;
    return (bytes % 4 == 0 && bytes / 4 <= 4) ? (int)(bytes / 4) : 0;
}

// grey_f32_kernel: a warp per 128 pixels (C = 3) or a thread per pixel (C = 4); uncapped
int grey_grid(const Img &d)
{
    size_t blocks = d.C == 3 ? (d.npix / 128 + 7) / 8 : (d.npix + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 1u << 20) blocks = 1u << 20;
    return (int)blocks;
}

int grid_for(int device, size_t work_items, int threads)
{
    size_t blocks = (work_items + threads - 1) / threads;
    size_t cap = (size_t)sm_count(device) * 8;  // 8 resident 256-thread CTAs per SM
    if (cap == 0) cap = 148 * 8;
    if (blocks > cap) blocks = cap;
    return blocks ? (int)blocks : 1;
}

// ---- Gaussian weights ------------------------------------------------------

// scipy.ndimage._gaussian_kernel1d: exp(-0.5/sigma^2 x^2), normalised in double
// over the nominal radius int(truncate*sigma + 0.5), truncate = 8 (the test
// oracle's setting, tests/millipyde_tests.py:550).
int oracle_weights(double sigma, double *w, int max_radius)
{
    int r = (int)(8.0 * sigma + 0.5);
    if (r > max_radius) {
        // Callers with a fixed-size table (the fusion pass's radius-bucket query): the taps beyond are
        // < exp(-32) of the centre.  launch_gaussian itself never truncates: larger radii take the
        // global-memory kernel with the full support (oracle_weights_full).
        r = max_radius;
    }
    double total = 0;
    for (int d = 0; d <= r; ++d) {
        w[d] = exp(-0.5 / (sigma * sigma) * (double)d * d);
        total += d ? 2 * w[d] : w[d];
    }
    for (int d = 0; d <= r; ++d) w[d] /= total;
    return r;
}

// The same kernel over its whole nominal support, whatever the radius.
std::vector<double> oracle_weights_full(double sigma)
{
    const int r = (int)(8.0 * sigma + 0.5);
    std::vector<double> w((size_t)r + 1);
    double total = 0;
    for (int d = 0; d <= r; ++d) {
        w[d] = exp(-0.5 / (sigma * sigma) * (double)d * d);
        total += d ? 2 * w[d] : w[d];
    }
    for (double &v : w) v /= total;
    return w;
}

// Smallest radius whose dropped tail mass is <= eps (weights already normalised).
int effective_radius(const double *w, int r, double eps)
{
    double tail = 0;
    int eff = r;
    for (int d = r; d > 0; --d) {
        tail += 2 * w[d];
        if (tail > eps) break;
        eff = d - 1;
    }
    return eff;
}

// The reference's 17 weights (src/millipyde_image.cpp:744-756): expf in float,
// normalised by the double sum taken in index order.
void reference_weights(double sigma, double *w_half)
{
    double full[17], total = 0;
    for (int i = 0; i < 17; ++i) {
        int dist = -1 * (8 - i);
        full[i] = expf(-1 * ((dist * dist) / (2 * sigma * sigma)));
        total += full[i];
    }
    for (int i = 0; i < 17; ++i) full[i] /= total;
    for (int d = 0; d <= 8; ++d) w_half[d] = full[8 + d];
}

}  // namespace mp

template <int K>
static void launch_transpose(const void *in, void *out, int W, int H, cudaStream_t s)
{
    dim3 grid((W + 31) / 32, (H + 31) / 32);
    if constexpr (K <= 2) {
        if (((size_t)W * K) % 4 == 0 && ((size_t)H * K) % 4 == 0) {
            constexpr int TX = 64 / K;
            transpose_tma64_kernel<K><<<dim3((W + TX - 1) / TX, (H + 63) / 64), 256, 0, s>>>(
                (const uint32_t *)in, (uint32_t *)out, W, H);
            return;
        }
    }
    if (((size_t)W * K) % 4 == 0 && ((size_t)H * K) % 4 == 0)
        transpose_tma_kernel<K, 64><<<dim3((W + 31) / 32, (H + 63) / 64), 256, 0, s>>>((const uint32_t *)in,
                                                                                      (uint32_t *)out, W, H);
    else
        transpose_kernel<K><<<grid, 256, 0, s>>>((const uint32_t *)in, (uint32_t *)out, W, H);
}

template <int K>
static void launch_fliplr(int dev, const void *in, void *out, int W, int H, cudaStream_t s)
{
    if (W % 128 == 0) {
        size_t nu = (size_t)H * (W / 128);
        size_t blocks = (nu + 7) / 8;
        if (blocks > (1u << 20)) blocks = 1u << 20;
        fliplr_warp_kernel<K><<<(unsigned)blocks, 256, 0, s>>>((const uint4 *)in, (uint4 *)out, W, nu);
    } else if (W % 4 == 0) {
        size_t nb = (size_t)H * (W / 4);
        fliplr_vec_kernel<K><<<mp::grid_for(dev, nb, 256), 256, 0, s>>>((const uint4 *)in, (uint4 *)out, W, nb);
    } else {
        size_t nw = (size_t)H * W * K;
        fliplr_scalar_kernel<K><<<mp::grid_for(dev, nw, 256), 256, 0, s>>>((const uint32_t *)in, (uint32_t *)out, W, nw);
    }
}

extern "C" {

void mpimg_set_semantics(int mode)
{
    g_semantics.store(mode == MP_SEMANTICS_REFERENCE ? MP_SEMANTICS_REFERENCE : MP_SEMANTICS_ORACLE);
}

int mpimg_get_semantics(void) { return semantics(); }

int mpimg_gaussian_effective_radius(double sigma, int *full)
{
    if (!(sigma > 1e-15)) {
        if (full) *full = 0;
        return 0;
    }
    double w[kGaussMaxRadius + 1];
    int r = mp::oracle_weights(sigma, w, kGaussMaxRadius);
    if (full) *full = (int)(8.0 * sigma + 0.5);
    return mp::effective_radius(w, r, mp::kGaussTailEps);
}

/* ------------------------------------------------------------------ rgb2grey */
static MPStatus grey_impl(MPObjData *obj, const PwProgram &pre, const PwProgram &post)
{
    mp::Img d;
    cudaStream_t s;
    MPStatus st = begin(obj, &d, &s);
    if (st != MILLIPYDE_SUCCESS) return st;
    if (obj->ndims != 3 || d.C < 3) return MP_ERROR_UNSUPPORTED_LAYOUT;

    const bool f32 = d.fam == mp::FAM_F32;
    const size_t out_es = f32 ? 4 : 8;
    const size_t out_bytes = d.npix * out_es;
    void *out;
    if ((st = fresh(obj, s, out_bytes, &out)) != MILLIPYDE_SUCCESS) return st;

    if (f32) {
        int grid = mp::grey_grid(d);
        if (d.C == 3)
            grey_f32_kernel<3><<<grid, 256, 0, s>>>((const float *)obj->device_data, (float *)out, d.npix, pre, post);
        else
            grey_f32_kernel<4><<<grid, 256, 0, s>>>((const float *)obj->device_data, (float *)out, d.npix, pre, post);
    } else if (d.fam == mp::FAM_RGBA8) {
        int grid = mp::grid_for(obj->mem_loc, d.npix / 4 + 1, 256);
        grey_rgba8_kernel<<<grid, 256, 0, s>>>((const uint32_t *)obj->device_data, (double *)out, d.npix);
    } else if (d.type == MP_NPY_UBYTE) {
        int grid = mp::grid_for(obj->mem_loc, d.npix, 256);
        grey_u8_generic_kernel<<<grid, 256, 0, s>>>((const uint8_t *)obj->device_data, (double *)out, d.npix, d.C);
    } else {  // fp64 colour
        int grid = mp::grid_for(obj->mem_loc, d.npix, 256);
        grey_f64_kernel<<<grid, 256, 0, s>>>((const double *)obj->device_data, (double *)out, d.npix, d.C);
    }
    mp::count_launch();
    if ((st = finish(obj, s, out, out_bytes)) != MILLIPYDE_SUCCESS) return st;

    // header rewrite, as src/millipyde_image.cpp:559-563 (type 12 = NPY_DOUBLE there)
    obj->ndims = 2;
    obj->type = f32 ? MP_NPY_FLOAT : MP_NPY_DOUBLE;
    obj->dims[2] = (int)(d.W * out_es);
    obj->dims[3] = (int)out_es;
    return MILLIPYDE_SUCCESS;
}

MPStatus mpimg_color_to_greyscale(MPObjData *obj, void *args)
{
    MP_UNUSED(args);
    PwProgram none = {};
    return grey_impl(obj, none, none);
}

/* ----------------------------------------------------------------- transpose */
MPStatus mpimg_transpose(MPObjData *obj, void *args)
{
    MP_UNUSED(args);
    mp::Img d;
    cudaStream_t s;
    MPStatus st = begin(obj, &d, &s);
    if (st != MILLIPYDE_SUCCESS) return st;
    int K = mp::words_per_pixel(d);
    if (!K) return MP_ERROR_UNSUPPORTED_LAYOUT;
    void *out;
    if ((st = fresh(obj, s, obj->nbytes, &out)) != MILLIPYDE_SUCCESS) return st;
    switch (K) {
        case 1: launch_transpose<1>(obj->device_data, out, d.W, d.H, s); break;
        case 2: launch_transpose<2>(obj->device_data, out, d.W, d.H, s); break;
        case 3: launch_transpose<3>(obj->device_data, out, d.W, d.H, s); break;
        default: launch_transpose<4>(obj->device_data, out, d.W, d.H, s); break;
    }
    mp::count_launch();
    if ((st = finish(obj, s, out, obj->nbytes)) != MILLIPYDE_SUCCESS) return st;
    // swap H and W; strides follow the new shape (the reference writes a wrong row
    // stride here, src/millipyde_image.cpp:888 -- harmless there, fixed here)
    const int pix_bytes = (int)(d.C * d.esize);
    obj->dims[0] = d.W;
    obj->dims[1] = d.H;
    obj->dims[obj->ndims] = d.H * pix_bytes;
    obj->dims[obj->ndims + 1] = pix_bytes;
    return MILLIPYDE_SUCCESS;
}

/* -------------------------------------------------------------------- fliplr */
MPStatus mpimg_fliplr(MPObjData *obj, void *args)
{
    MP_UNUSED(args);
    mp::Img d;
    cudaStream_t s;
    MPStatus st = begin(obj, &d, &s);
    if (st != MILLIPYDE_SUCCESS) return st;
    int K = mp::words_per_pixel(d);
    if (!K) return MP_ERROR_UNSUPPORTED_LAYOUT;
    void *out;
    if ((st = fresh(obj, s, obj->nbytes, &out)) != MILLIPYDE_SUCCESS) return st;
    switch (K) {
        case 1: launch_fliplr<1>(obj->mem_loc, obj->device_data, out, d.W, d.H, s); break;
        case 2: launch_fliplr<2>(obj->mem_loc, obj->device_data, out, d.W, d.H, s); break;
        case 3: launch_fliplr<3>(obj->mem_loc, obj->device_data, out, d.W, d.H, s); break;
        default: launch_fliplr<4>(obj->mem_loc, obj->device_data, out, d.W, d.H, s); break;
    }
    mp::count_launch();
    return finish(obj, s, out, obj->nbytes);
}

/* -------------------------------------------------------------------- rotate */
MPStatus mpimg_rotate(MPObjData *obj, void *args)
{
    if (!args) return MP_ERROR_INVALID_ARGUMENT;
    const double angle = ((RotateArgs *)args)->angle;
    mp::Img d;
    cudaStream_t s;
    MPStatus st = begin(obj, &d, &s);
    if (st != MILLIPYDE_SUCCESS) return st;
    if (d.fam != mp::FAM_F32 && d.fam != mp::FAM_F64 && d.fam != mp::FAM_RGBA8) return MP_ERROR_UNSUPPORTED_LAYOUT;

    void *out;
    if ((st = fresh(obj, s, obj->nbytes, &out)) != MILLIPYDE_SUCCESS) return st;

    const bool nearest = d.fam == mp::FAM_RGBA8 || (d.fam == mp::FAM_F64 && semantics() == MP_SEMANTICS_REFERENCE);
    // reference layouts: through a staged box when the rows are TMA-copyable (16-byte aligned), else the direct kernels
    const int K = d.fam == mp::FAM_RGBA8 ? 1 : 2;
    const bool box_ok = d.fam != mp::FAM_F32 && ((size_t)d.W * K) % 4 == 0 &&
                        (reinterpret_cast<uintptr_t>(obj->device_data) & 15) == 0;
    if (nearest) {
        const double rad = angle * 0.01745329252;  // src/millipyde_image.cpp:702
        RotateParams none = {};
        if (box_ok && K == 1) {
            dim3 grid((d.W + 31) / 32, (d.H + RotBoxGeom<1>::TH - 1) / RotBoxGeom<1>::TH);
            rotate_box_kernel<1, false><<<grid, 256, RotBoxGeom<1>::SMEM, s>>>((const uint32_t *)obj->device_data,
                                                                               (uint32_t *)out, d.W, d.H, rad, none);
        } else if (box_ok) {
            dim3 grid((d.W + 31) / 32, (d.H + RotBoxGeom<2>::TH - 1) / RotBoxGeom<2>::TH);
            rotate_box_kernel<2, false><<<grid, 256, RotBoxGeom<2>::SMEM, s>>>((const uint32_t *)obj->device_data,
                                                                               (uint32_t *)out, d.W, d.H, rad, none);
        } else {
            dim3 block(32, 8), grid((d.W + 31) / 32, (d.H + 7) / 8);
            if (d.fam == mp::FAM_RGBA8)
                rotate_nearest_kernel<1><<<grid, block, 0, s>>>((const uint32_t *)obj->device_data, (uint32_t *)out, d.W, d.H, rad);
            else
                rotate_nearest_kernel<2><<<grid, block, 0, s>>>((const uint32_t *)obj->device_data, (uint32_t *)out, d.W, d.H, rad);
        }
    } else {
        RotateParams rp = mp::rotate_params(d.W, d.H, angle);
        if (d.fam == mp::FAM_F64 && box_ok) {
            dim3 grid((d.W + 31) / 32, (d.H + RotBoxGeom<2>::TH - 1) / RotBoxGeom<2>::TH);
            rotate_box_kernel<2, true><<<grid, 256, RotBoxGeom<2>::SMEM, s>>>((const uint32_t *)obj->device_data,
                                                                              (uint32_t *)out, d.W, d.H, 0.0, rp);
        } else if (d.fam == mp::FAM_F64) {
            dim3 grid((d.W + 31) / 32, (d.H + 7) / 8);
            rotate_bilinear_kernel<double, 1><<<grid, 256, 0, s>>>((const double *)obj->device_data, (double *)out, d.W, d.H, rp);
        } else {
            // fp32: the tile-staged gather (kernels/geometry.cuh) with identity index maps and no
            // pointwise programs -- its loads are coalesced at every angle
            GatherParams g = {};
            g.in = (const float *)obj->device_data;
            g.out = (float *)out;
            g.out_h = g.rot_h = d.H;
            g.out_w = g.rot_w = g.src_w = d.W;
            g.post = g.pre = IndexMap{1, 0, 0, 0, 1, 0};
            g.has_rotate = 1;
            g.rp = rp;
            mp::launch_gather_f32(s, d.C, g, 1);
            return finish(obj, s, out, obj->nbytes);
        }
    }
    mp::count_launch();
    return finish(obj, s, out, obj->nbytes);
}

/* ------------------------------------------------ brightness / gamma / colorize */
static MPStatus pointwise_f32(MPObjData *obj, const mp::Img &d, cudaStream_t s, const PwProgram &prog)
{
    if (prog.n == 0) return MILLIPYDE_SUCCESS;
    void *out;
    MPStatus st = fresh(obj, s, obj->nbytes, &out);
    if (st != MILLIPYDE_SUCCESS) return st;
    size_t n = d.npix * d.C;
    // one warp per 96 vectors, no cap: many small CTAs stream better than a short persistent grid
    size_t want = (n / 4 + 96 * 8 - 1) / (96 * 8);
    int grid = (int)(want < 1 ? 1 : (want > 65535u * 16u ? 65535u * 16u : want));
    const float *in = (const float *)obj->device_data;
    if (d.C == 1) pw_f32_kernel<1><<<grid, 256, 0, s>>>(in, (float *)out, n, prog);
    else if (d.C == 3) pw_f32_kernel<3><<<grid, 256, 0, s>>>(in, (float *)out, n, prog);
    else pw_f32_kernel<4><<<grid, 256, 0, s>>>(in, (float *)out, n, prog);
    mp::count_launch();
    return finish(obj, s, out, obj->nbytes);
}

static MPStatus pointwise_f64(MPObjData *obj, const mp::Img &d, cudaStream_t s, PwOp64 op)
{
    void *out;
    MPStatus st = fresh(obj, s, obj->nbytes, &out);
    if (st != MILLIPYDE_SUCCESS) return st;
    size_t n = d.npix;
    pw_f64_kernel<<<mp::grid_for(obj->mem_loc, n / 2 + 1, 256), 256, 0, s>>>(
        (const double *)obj->device_data, (double *)out, n, op, semantics() == MP_SEMANTICS_REFERENCE);
    mp::count_launch();
    return finish(obj, s, out, obj->nbytes);
}

// Device-resident byte tables per (device, program), built on first use and kept: programs are few
// (an Operation's arguments) and a table is 768 bytes.  Beyond kLutCacheMax distinct programs
// (e.g. a stream of random_* draws) tables are built into a stream-ordered temporary instead.
static const uint8_t *u8_lut_for(int device, cudaStream_t s, const U8Program &prog, void **temp)
{
    static std::mutex mux;
    static std::map<std::string, const uint8_t *> cache[64];
    constexpr size_t kLutCacheMax = 256;
    *temp = nullptr;
    std::string key((const char *)&prog, sizeof prog);
    if (device >= 0 && device < 64) {
        std::lock_guard<std::mutex> lk(mux);
        auto it = cache[device].find(key);
        if (it != cache[device].end()) return it->second;
        if (cache[device].size() < kLutCacheMax) {
            uint8_t *lut = nullptr;
            if (cudaMalloc((void **)&lut, 768) == cudaSuccess) {
                u8_lut_kernel<<<3, 256, 0, s>>>(lut, prog);
                mp::count_launch();
                cudaStreamSynchronize(s);  // once per program: other streams may use the table next
                cache[device][key] = lut;
                return lut;
            }
            (void)cudaGetLastError();
        }
    }
    uint8_t *lut = (uint8_t *)mp::pool_alloc(device, s, 768);
    if (!lut) return nullptr;
    u8_lut_kernel<<<3, 256, 0, s>>>(lut, prog);
    mp::count_launch();
    *temp = lut;
    return lut;
}

static MPStatus pointwise_rgba8(MPObjData *obj, const mp::Img &d, cudaStream_t s, const U8Program &prog)
{
    if (prog.n == 0) return MILLIPYDE_SUCCESS;
    void *out;
    MPStatus st = fresh(obj, s, obj->nbytes, &out);
    if (st != MILLIPYDE_SUCCESS) return st;
    void *temp = nullptr;
    const uint8_t *lut = u8_lut_for(obj->mem_loc, s, prog, &temp);
    if (!lut) {
        mp::pool_free(obj->mem_loc, s, out);
        return MP_ERROR_DEVICE_ALLOC;
    }
    const size_t ngroups = d.npix / 4;
    const size_t blocks = (ngroups + 256 * kU8PwVecs - 1) / (256 * kU8PwVecs);
    pw_rgba8_kernel<<<(unsigned)(blocks ? blocks : 1), 256, 0, s>>>((const uint32_t *)obj->device_data, (uint32_t *)out,
                                                                d.npix, lut);
    mp::count_launch();
    if (temp) mp::pool_free(obj->mem_loc, s, temp);
    return finish(obj, s, out, obj->nbytes);
}

MPStatus mpimg_brightness(MPObjData *obj, void *args)
{
    if (!args) return MP_ERROR_INVALID_ARGUMENT;
    const double delta = ((BrightnessArgs *)args)->delta;
    mp::Img d;
    cudaStream_t s;
    MPStatus st = begin(obj, &d, &s);
    if (st != MILLIPYDE_SUCCESS) return st;
    if (d.fam == mp::FAM_F32) {
        PwProgram p = {};
        p.n = 1;
        p.ops[0] = PwOp{PW_BRIGHTNESS, (float)delta, 0.f, 0.f};
        return pointwise_f32(obj, d, s, p);
    }
    if (d.fam == mp::FAM_F64) return pointwise_f64(obj, d, s, PwOp64{PW_BRIGHTNESS, delta, 0});
    if (d.fam == mp::FAM_RGBA8) {
        U8Program p = {};
        p.n = 1;
        p.ops[0] = mp::u8_brightness_op(delta);
        return pointwise_rgba8(obj, d, s, p);
    }
    return MP_ERROR_UNSUPPORTED_LAYOUT;
}

MPStatus mpimg_elementwise(MPObjData *obj, void *args)
{
    if (!args) return MP_ERROR_INVALID_ARGUMENT;
    mp::Img d;
    cudaStream_t s;
    MPStatus st = begin(obj, &d, &s);
    if (st != MILLIPYDE_SUCCESS) return st;
    if (d.fam != mp::FAM_F32) return MP_ERROR_UNSUPPORTED_LAYOUT;
    PwProgram p = {};
    st = mp::op_elementwise_program((const ElementwiseArgs *)args, d.C, &p);
    if (st != MILLIPYDE_SUCCESS) return st;
    return pointwise_f32(obj, d, s, p);
}

MPStatus mpimg_adjust_gamma(MPObjData *obj, void *args)
{
    if (!args) return MP_ERROR_INVALID_ARGUMENT;
    const double gamma = ((GammaArgs *)args)->gamma, gain = ((GammaArgs *)args)->gain;
    mp::Img d;
    cudaStream_t s;
    MPStatus st = begin(obj, &d, &s);
    if (st != MILLIPYDE_SUCCESS) return st;
    if (d.fam == mp::FAM_F32) {
        PwProgram p = {};
        p.n = 1;
        p.ops[0] = PwOp{PW_GAMMA, (float)gamma, (float)gain, 0.f};
        return pointwise_f32(obj, d, s, p);
    }
    if (d.fam == mp::FAM_F64) return pointwise_f64(obj, d, s, PwOp64{PW_GAMMA, gamma, gain});
    if (d.fam == mp::FAM_RGBA8) {
        U8Program p = {};
        p.n = 1;
        p.ops[0] = U8Op{PW_GAMMA, 0, gamma, gain, 0};
        return pointwise_rgba8(obj, d, s, p);
    }
    return MP_ERROR_UNSUPPORTED_LAYOUT;
}

MPStatus mpimg_colorize(MPObjData *obj, void *args)
{
    if (!args) return MP_ERROR_INVALID_ARGUMENT;
    const ColorizeArgs *a = (const ColorizeArgs *)args;
    mp::Img d;
    cudaStream_t s;
    MPStatus st = begin(obj, &d, &s);
    if (st != MILLIPYDE_SUCCESS) return st;
    if (d.C == 1) return MILLIPYDE_SUCCESS;  // no colorization for grey images (:647-651)
    if (d.fam == mp::FAM_F32) {
        PwProgram p = {};
        p.n = 1;
        p.ops[0] = PwOp{PW_COLORIZE, (float)a->r_mult, (float)a->g_mult, (float)a->b_mult};
        return pointwise_f32(obj, d, s, p);
    }
    if (d.fam == mp::FAM_RGBA8) {
        U8Program p = {};
        p.n = 1;
        p.ops[0] = U8Op{PW_COLORIZE, 0, a->r_mult, a->g_mult, a->b_mult};
        return pointwise_rgba8(obj, d, s, p);
    }
    return MP_ERROR_UNSUPPORTED_LAYOUT;
}

/* ------------------------------------------------------------------ gaussian */
MPStatus mpimg_gaussian(MPObjData *obj, void *args)
{
    if (!args) return MP_ERROR_INVALID_ARGUMENT;
    const double sigma = ((GaussianArgs *)args)->sigma;
    mp::Img d;
    cudaStream_t s;
    MPStatus st = begin(obj, &d, &s);
    if (st != MILLIPYDE_SUCCESS) return st;
    if (d.fam != mp::FAM_F32 && d.fam != mp::FAM_F64 && d.fam != mp::FAM_RGBA8) return MP_ERROR_UNSUPPORTED_LAYOUT;
    if (!(sigma >= 0)) return MP_ERROR_INVALID_ARGUMENT;

    const bool ref_rule = d.fam == mp::FAM_RGBA8 || (d.fam == mp::FAM_F64 && semantics() == MP_SEMANTICS_REFERENCE);
    if (!ref_rule && !(sigma > 1e-15)) return MILLIPYDE_SUCCESS;  // scipy: sigma ~ 0 is a copy

    void *out;
    if ((st = fresh(obj, s, obj->nbytes, &out)) != MILLIPYDE_SUCCESS) return st;
    st = mp::launch_gaussian(obj->mem_loc, s, d, obj->device_data, out, sigma, ref_rule);
    if (st != MILLIPYDE_SUCCESS) {
        mp::pool_free(obj->mem_loc, s, out);
        return st;
    }
    return finish(obj, s, out, obj->nbytes);
}

/* ------------------------------------------------------------------ random_* */
MPStatus mpimg_random_rotate(MPObjData *obj, void *args)
{
    if (!args) return MP_ERROR_INVALID_ARGUMENT;
    const RandomRangeArgs *r = (const RandomRangeArgs *)args;
    RotateArgs a;
    MPStatus st = random_double_in_range(r->min, r->max, &a.angle);
    return st != MILLIPYDE_SUCCESS ? st : mpimg_rotate(obj, &a);
}

MPStatus mpimg_random_gaussian(MPObjData *obj, void *args)
{
    if (!args) return MP_ERROR_INVALID_ARGUMENT;
    const RandomRangeArgs *r = (const RandomRangeArgs *)args;
    GaussianArgs a;
    MPStatus st = random_double_in_range(r->min, r->max, &a.sigma);
    return st != MILLIPYDE_SUCCESS ? st : mpimg_gaussian(obj, &a);
}

MPStatus mpimg_random_brightness(MPObjData *obj, void *args)
{
    if (!args) return MP_ERROR_INVALID_ARGUMENT;
    const RandomRangeArgs *r = (const RandomRangeArgs *)args;
    BrightnessArgs a;
    MPStatus st = random_double_in_range(r->min, r->max, &a.delta);
    return st != MILLIPYDE_SUCCESS ? st : mpimg_brightness(obj, &a);
}

MPStatus mpimg_random_adjust_gamma(MPObjData *obj, void *args)
{
    if (!args) return MP_ERROR_INVALID_ARGUMENT;
    const RandomGammaArgs *r = (const RandomGammaArgs *)args;
    GammaArgs a;
    MPStatus st = random_double_in_range(r->gamma_min, r->gamma_max, &a.gamma);
    if (st == MILLIPYDE_SUCCESS) st = random_double_in_range(r->gain_min, r->gain_max, &a.gain);
    return st != MILLIPYDE_SUCCESS ? st : mpimg_adjust_gamma(obj, &a);
}

MPStatus mpimg_random_colorize(MPObjData *obj, void *args)
{
    if (!args) return MP_ERROR_INVALID_ARGUMENT;
    const RandomColorizeArgs *r = (const RandomColorizeArgs *)args;
    ColorizeArgs a;
    MPStatus st = random_double_in_range(r->r_min, r->r_max, &a.r_mult);
    if (st == MILLIPYDE_SUCCESS) st = random_double_in_range(r->g_min, r->g_max, &a.g_mult);
    if (st == MILLIPYDE_SUCCESS) st = random_double_in_range(r->b_min, r->b_max, &a.b_mult);
    return st != MILLIPYDE_SUCCESS ? st : mpimg_colorize(obj, &a);
}

MPFunc mpimg_func_from_name(const char *name, size_t *arg_bytes)
{
    static const struct {
        const char *name;
        MPFunc func;
        size_t bytes;
    } table[] = {
        {"rgb2grey", mpimg_color_to_greyscale, 0},
        {"rgb2gray", mpimg_color_to_greyscale, 0},
        {"rgba2grey", mpimg_color_to_greyscale, 0},
        {"rgba2gray", mpimg_color_to_greyscale, 0},
        {"transpose", mpimg_transpose, 0},
        {"fliplr", mpimg_fliplr, 0},
        {"gaussian", mpimg_gaussian, sizeof(GaussianArgs)},
        {"rotate", mpimg_rotate, sizeof(RotateArgs)},
        {"brightness", mpimg_brightness, sizeof(BrightnessArgs)},
        {"adjust_gamma", mpimg_adjust_gamma, sizeof(GammaArgs)},
        {"colorize", mpimg_colorize, sizeof(ColorizeArgs)},
        {"random_rotate", mpimg_random_rotate, sizeof(RandomRangeArgs)},
        {"random_gaussian", mpimg_random_gaussian, sizeof(RandomRangeArgs)},
        {"random_brightness", mpimg_random_brightness, sizeof(RandomRangeArgs)},
        {"random_adjust_gamma", mpimg_random_adjust_gamma, sizeof(RandomGammaArgs)},
        {"random_colorize", mpimg_random_colorize, sizeof(RandomColorizeArgs)},
    };
    if (arg_bytes) *arg_bytes = 0;
    if (!name) return NULL;
    for (size_t i = 0; i < sizeof(table) / sizeof(table[0]); ++i) {
        if (strcmp(name, table[i].name) == 0) {
            if (arg_bytes) *arg_bytes = table[i].bytes;
            return table[i].func;
        }
    }
    return NULL;
}

}  // extern "C"

namespace mp {

RotateParams rotate_params(int W, int H, double angle_deg)
{
    const double t = angle_deg * (M_PI / 180.0);  // np.deg2rad
    RotateParams rp;
    rp.c = cos(t);
    rp.s = sin(t);
    rp.cx = W / 2.0 - 0.5;
    rp.cy = H / 2.0 - 0.5;
    return rp;
}

U8Op u8_brightness_op(double delta)
{
    // char delta_n = (char)(delta * 255), on the host (src/millipyde_image.cpp:613)
    return U8Op{PW_BRIGHTNESS, (int)(signed char)(delta * 255), 0, 0, 0};
}

template <typename T, int C, bool CLAMP0>
static MPStatus launch_tile(cudaStream_t s, const Img &d, const void *in, void *out, const GaussParams<T> &gp)
{
    const int R = gp.radius;
    auto bytes_for = [&](int t) {
        return ((size_t)(t + 2 * R) * (t + 2 * R) * C + (size_t)(t + 2 * R) * t * C) * sizeof(T);
    };
    int tile = 32;
    while (tile > 8 && bytes_for(tile) > 200 * 1024) tile /= 2;
    size_t smem = bytes_for(tile);
    if (smem > 220 * 1024) return MP_ERROR_INVALID_ARGUMENT;  // sigma beyond what one tile can hold
    auto kern = gauss_tile_kernel<T, C, CLAMP0>;
    MP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((d.W + tile - 1) / tile, (d.H + tile - 1) / tile);
    kern<<<grid, 256, smem, s>>>((const T *)in, (T *)out, d.W, d.H, tile, tile, gp);
    count_launch();
    return MILLIPYDE_SUCCESS;
}

// fp64 greyscale with a compiled radius (8 or 16); false if the radius has no instance.
template <int R, bool CLAMP0>
static void launch_f64_fixed(int device, cudaStream_t s, const Img &d, const void *in, void *out,
                             const GaussParams<double> &gp)
{
    using G = GaussF64Geom<R>;
    static std::atomic<bool> configured[64] = {};  // the attribute is per device (and instantiation); worker threads race here
    if (device >= 0 && device < 64 && !configured[device].load(std::memory_order_acquire)) {
        cudaFuncSetAttribute(gauss_f64_kernel<R, CLAMP0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
        configured[device].store(true, std::memory_order_release);
    }
    dim3 grid((d.W + G::TW - 1) / G::TW, (d.H + G::TH - 1) / G::TH);
    // TMA staging needs rows of whole 16-byte vectors (even width, 16-byte-aligned base)
    const int tma = (d.W % 2 == 0) && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    gauss_f64_kernel<R, CLAMP0><<<grid, 256, G::SMEM, s>>>((const double *)in, (double *)out, d.W, d.H, gp, tma);
    count_launch();
}

// Any radius, any float type: rows then columns through a pool-allocated intermediate
// (kernels/gaussian_global.cuh).  `w` holds radius + 1 weights on the host.
template <typename T>
static MPStatus launch_global(int device, cudaStream_t s, const Img &d, const void *in, void *out, const double *w,
                              int radius)
{
    const size_t total = d.npix * (size_t)d.C;
    T *tmp = (T *)pool_alloc(device, s, total * sizeof(T));
    double *dw = (double *)pool_alloc(device, s, ((size_t)radius + 1) * sizeof(double));
    if (!tmp || !dw) {
        if (tmp) pool_free(device, s, tmp);
        if (dw) pool_free(device, s, dw);
        return MP_ERROR_DEVICE_ALLOC;
    }
    // pageable source: the runtime stages it before the call returns, `w` may die afterwards
    MP_CUDA_TRY(cudaMemcpyAsync(dw, w, ((size_t)radius + 1) * sizeof(double), cudaMemcpyHostToDevice, s));
    const int grid = grid_for(device, total, 256);
    gauss_global_pass_kernel<T, true><<<grid, 256, 0, s>>>((const T *)in, tmp, total, d.W * d.C, d.C, d.W, dw, radius);
    gauss_global_pass_kernel<T, false><<<grid, 256, 0, s>>>(tmp, (T *)out, total, d.W * d.C, d.C, d.H, dw, radius);
    count_launch(2);
    pool_free(device, s, tmp);
    pool_free(device, s, dw);
    return MILLIPYDE_SUCCESS;
}

MPStatus launch_gaussian(int device, cudaStream_t s, const Img &d, const void *in, void *out, double sigma,
                         bool ref_rule)
{
    double w[kGaussMaxRadius + 1];
    if (d.fam == FAM_RGBA8) {
        GaussParams<double> gp = {};
        gp.radius = 8;
        reference_weights(sigma, gp.w);
        const int R = 8, tile = 32;
        size_t smem = ((size_t)(tile + 2 * R) * (tile + 2 * R) + (size_t)(tile + 2 * R) * tile) * 4;
        dim3 grid((d.W + tile - 1) / tile, (d.H + tile - 1) / tile);
        // exact integer form, if every (byte, weight) product agrees with the double rule
        GaussU8Params ip = {};
        ip.radius = R;
        bool exact = true;
        for (int k = 0; k <= R && exact; ++k) {
            const double w = gp.w[k];
            bool found = false;
            const double scaled = w * 4294967296.0;
            for (int bump = 0; bump <= 1 && !found; ++bump) {
                if (!(scaled >= 0 && scaled < 4294967295.0)) break;
                const uint32_t m = (uint32_t)scaled + (uint32_t)bump;
                bool ok = true;
                for (uint32_t b = 0; b < 256 && ok; ++b)
                    ok = (uint32_t)(((uint64_t)b * m) >> 32) == (uint32_t)(int)(b * w);
                if (ok) {
                    ip.m[k] = m;
                    found = true;
                }
            }
            exact = found;
        }
        // fp32 chain form (one FFMA2.RM per tap and channel pair), if floor(b * (float)w) agrees with
        // the double rule for every byte and tap; b * wf is exact in double (8 x 24 bits)
        GaussU8ChainParams cp = {};
        bool chain_ok = true;
        for (int k = 0; k <= R && chain_ok; ++k) {
            bool found = false;
            for (int nudge = 0; nudge < 3 && !found; ++nudge) {
                float wf = (float)gp.w[k];
                if (nudge == 1) wf = nextafterf(wf, 1.f);
                if (nudge == 2) wf = nextafterf(wf, 0.f);
                bool ok = true;
                for (int b = 0; b < 256 && ok; ++b) ok = (int)floor((double)b * (double)wf) == (int)(b * gp.w[k]);
                if (ok) {
                    uint32_t bits;
                    memcpy(&bits, &wf, 4);
                    cp.ww[k] = ((unsigned long long)bits << 32) | bits;
                    found = true;
                }
            }
            chain_ok = found;
        }
        if (chain_ok) {
            // taps with 255 * w < 1 add (int)(byte * w) = 0 for every byte: not evaluated (gaussian_tile.cuh)
            int r_eff = 1;
            for (int k = 1; k <= R; ++k)
                if ((int)(255 * gp.w[k]) >= 1) r_eff = k;
            dim3 cgrid((d.W + kU8TW - 1) / kU8TW, (d.H + kU8TH - 1) / kU8TH);
            auto go = [&](auto rc) {
                constexpr int RR = decltype(rc)::value;
                static std::atomic<bool> configured[64] = {};  // the attribute is per device (and instantiation)
                if (device >= 0 && device < 64 && !configured[device].load(std::memory_order_acquire)) {
                    cudaFuncSetAttribute(gauss_rgba8_chain_kernel<RR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)U8Geom<RR>::Smem);
                    configured[device].store(true, std::memory_order_release);
                }
                gauss_rgba8_chain_kernel<RR><<<cgrid, 256, U8Geom<RR>::Smem, s>>>((const uint32_t *)in, (uint32_t *)out,
                                                                                 d.W, d.H, cp);
            };
            if (r_eff <= 3) go(std::integral_constant<int, 3>{});
            else if (r_eff <= 4) go(std::integral_constant<int, 4>{});
            else if (r_eff <= 5) go(std::integral_constant<int, 5>{});
            else if (r_eff <= 6) go(std::integral_constant<int, 6>{});
            else go(std::integral_constant<int, 8>{});
            count_launch();
            return MILLIPYDE_SUCCESS;
        }
        if (exact) {
            gauss_rgba8_int_kernel<<<grid, 256, smem, s>>>((const uint32_t *)in, (uint32_t *)out, d.W, d.H, tile, tile, ip);
            count_launch();
            return MILLIPYDE_SUCCESS;
        }
        gauss_rgba8_tile_kernel<<<grid, 256, smem, s>>>((const uint32_t *)in, (uint32_t *)out, d.W, d.H, tile, tile, gp);
        count_launch();
        return MILLIPYDE_SUCCESS;
    }
    if (d.fam == FAM_F64) {
        GaussParams<double> gp = {};
        if (ref_rule) {
            gp.radius = 8;
            reference_weights(sigma, gp.w);
            launch_f64_fixed<8, true>(device, s, d, in, out, gp);
            return MILLIPYDE_SUCCESS;
        }
        const int nominal = (int)(8.0 * sigma + 0.5);
        if (nominal > kGaussMaxRadius) {   // sigma > ~15.9: the full support, no truncation (matches scipy to 1e-12)
            const std::vector<double> wf = oracle_weights_full(sigma);
            return launch_global<double>(device, s, d, in, out, wf.data(), nominal);
        }
        gp.radius = oracle_weights(sigma, gp.w, kGaussMaxRadius);
        if (gp.radius == 16) {  // sigma = 2 under the oracle rule (truncate = 8)
            launch_f64_fixed<16, false>(device, s, d, in, out, gp);
            return MILLIPYDE_SUCCESS;
        }
        if (gp.radius == 8) {
            launch_f64_fixed<8, false>(device, s, d, in, out, gp);
            return MILLIPYDE_SUCCESS;
        }
        {
            const MPStatus st = launch_tile<double, 1, false>(s, d, in, out, gp);
            if (st != MP_ERROR_INVALID_ARGUMENT) return st;
            return launch_global<double>(device, s, d, in, out, gp.w, gp.radius);   // the tile + halo exceeds shared memory
        }
    }
    // fp32: scipy weights, evaluated over the effective support only
    if ((int)(8.0 * sigma + 0.5) > kGaussMaxRadius) {
        const std::vector<double> wf = oracle_weights_full(sigma);
        const int eff_full = effective_radius(wf.data(), (int)wf.size() - 1, mp::kGaussTailEps);
        return launch_global<float>(device, s, d, in, out, wf.data(), eff_full);
    }
    int r = oracle_weights(sigma, w, kGaussMaxRadius);
    int eff = effective_radius(w, r, mp::kGaussTailEps);
    GaussParams<float> gp = {};
    gp.radius = eff;
    for (int k = 0; k <= eff; ++k) gp.w[k] = (float)w[k];
    if (gauss_stream_supported(d.W, d.C, eff))
        return launch_gauss_stream(device, s, d, (const float *)in, (float *)out, gp);
    MPStatus st;
    if (d.C == 1) st = launch_tile<float, 1, false>(s, d, in, out, gp);
    else if (d.C == 3) st = launch_tile<float, 3, false>(s, d, in, out, gp);
    else st = launch_tile<float, 4, false>(s, d, in, out, gp);
    if (st != MP_ERROR_INVALID_ARGUMENT) return st;
    return launch_global<float>(device, s, d, in, out, w, eff);   // the tile + halo exceeds shared memory
}

}  // namespace mp

namespace mp {

MPStatus op_pointwise_f32(MPObjData *obj, const PwProgram &prog)
{
    Img d;
    cudaStream_t s;
    MPStatus st = begin(obj, &d, &s);
    if (st != MILLIPYDE_SUCCESS) return st;
    if (d.fam != FAM_F32) return MP_ERROR_UNSUPPORTED_LAYOUT;
    return pointwise_f32(obj, d, s, prog);
}

MPStatus op_elementwise_program(const ElementwiseArgs *a, int channels, PwProgram *prog)
{
    const int kind = (int)a->kind;
    if (kind < MP_EW_ADD || kind > MP_EW_CLIP) return MP_ERROR_INVALID_ARGUMENT;
    PwOp op = {kind, (float)a->a, (float)a->b, (float)a->c};
    if (kind == MP_EW_MUL) {
        if (a->per_channel != 0) {
            if (channels != 3) return MP_ERROR_UNSUPPORTED_LAYOUT;
        } else {
            op.b = op.c = op.a;
        }
    }
    prog->n = 1;
    prog->ops[0] = op;
    return MILLIPYDE_SUCCESS;
}

MPStatus op_pointwise_rgba8(MPObjData *obj, const U8Program &prog)
{
    Img d;
    cudaStream_t s;
    MPStatus st = begin(obj, &d, &s);
    if (st != MILLIPYDE_SUCCESS) return st;
    if (d.fam != FAM_RGBA8) return MP_ERROR_UNSUPPORTED_LAYOUT;
    return pointwise_rgba8(obj, d, s, prog);
}

MPStatus op_grey_f32(MPObjData *obj, const PwProgram &pre, const PwProgram &post)
{
    return grey_impl(obj, pre, post);
}

}  // namespace mp

namespace mp {

// Row pitch of the gather's staged box for one rotation.  A warp reads the corners of 32 consecutive
// output pixels, i.e. a line in the source at the rotation's angle: the word address of lane l is
// C * ix(l) + pitch * iy(l) (+ channel, + corner), so how many lanes collide on a bank depends on the
// angle and on pitch mod 32.  The launcher evaluates the few legal pitches (multiples of 4: the staged
// rows are TMA destinations) on a handful of sub-pixel offsets and takes the best; at 30 degrees and
// C = 3 the default pitch serialises 3 ways, the best 2 (the floor for mid angles).
template <int C, int TH>
static int gather_pitch_for(const RotateParams &rp)
{
    using G = GatherGeom<C, TH>;
    int best_pitch = G::PITCH;
    long best = -1;
    for (int pitch = G::PITCH; pitch <= G::PITCH_MAX; pitch += 4) {
        long cost = 0;
        for (int sample = 0; sample < 12; ++sample) {
            const double x0 = 40.0 + 0.37 * sample + 0.11 * (sample % 5), y0 = 40.0 + 0.53 * sample + 0.07 * (sample % 3);
            for (int corner = 0; corner < 4; ++corner)
                for (int c = 0; c < C; ++c) {
                    // distinct addresses per bank, maximum over the banks = wavefronts of this load
                    long addr[32];
                    for (int l = 0; l < 32; ++l) {
                        const int ix = (int)floor(x0 + rp.c * l) + (corner & 1), iy = (int)floor(y0 + rp.s * l) + (corner >> 1);
                        addr[l] = (long)iy * pitch + (long)ix * C + c;
                    }
                    int worst = 0;
                    for (int b = 0; b < 32; ++b) {
                        int distinct = 0;
                        long seen[32];
                        for (int l = 0; l < 32; ++l) {
                            if (((addr[l] % 32) + 32) % 32 != b) continue;
                            bool dup = false;
                            for (int k = 0; k < distinct; ++k) dup = dup || seen[k] == addr[l];
                            if (!dup) seen[distinct++] = addr[l];
                        }
                        worst = distinct > worst ? distinct : worst;
                    }
                    cost += worst;
                }
        }
        if (best < 0 || cost < best) {
            best = cost;
            best_pitch = pitch;
        }
    }
    return best_pitch;
}

void launch_gather_f32(cudaStream_t s, int channels, const GatherParams &g_in, int n_images)
{
    GatherParams g = g_in;
    g.pitch = 0;
    if (g.has_rotate && !g.var_tab) {   // one angle for the whole launch: lay the box out for it
        static thread_local double memo_c[3] = {2, 2, 2}, memo_s[3] = {2, 2, 2};
        static thread_local int memo_pitch[3] = {0, 0, 0};
        const int slot = channels == 1 ? 0 : (channels == 3 ? 1 : 2);
        if (memo_c[slot] != g.rp.c || memo_s[slot] != g.rp.s) {
            memo_pitch[slot] = channels == 1 ? gather_pitch_for<1, kGatherTileTall>(g.rp)
                                             : (channels == 3 ? gather_pitch_for<3, kGatherTile>(g.rp)
                                                              : gather_pitch_for<4, kGatherTile>(g.rp));
            memo_c[slot] = g.rp.c;
            memo_s[slot] = g.rp.s;
        }
        g.pitch = memo_pitch[slot];
    }
    // single-channel images use 32 x 64 tiles: a 32 x 32 tile of 4-byte pixels is too little work per CTA
    const int th = channels == 1 ? kGatherTileTall : kGatherTile;
    dim3 grid((g.out_w + kGatherTile - 1) / kGatherTile, (g.out_h + th - 1) / th, n_images);
    const size_t smem = channels == 1 ? GatherGeom<1, kGatherTileTall>::SMEM
                                      : (channels == 3 ? GatherGeom<3>::SMEM : GatherGeom<4>::SMEM);
    static_assert(GatherGeom<4>::SMEM <= 48 * 1024 && GatherGeom<1, kGatherTileTall>::SMEM <= 48 * 1024,
                  "fits the default dynamic shared memory limit on every device");
    if (g.var_tab) {  // per-image angle / programs
        if (channels == 1) gather_f32_kernel<1, true, kGatherTileTall><<<grid, 256, smem, s>>>(g);
        else if (channels == 3) gather_f32_kernel<3, true><<<grid, 256, smem, s>>>(g);
        else gather_f32_kernel<4, true><<<grid, 256, smem, s>>>(g);
    } else if (channels == 1) gather_f32_kernel<1, false, kGatherTileTall><<<grid, 256, smem, s>>>(g);
    else if (channels == 3) gather_f32_kernel<3><<<grid, 256, smem, s>>>(g);
    else gather_f32_kernel<4><<<grid, 256, smem, s>>>(g);
    count_launch();
}

}  // namespace mp

namespace mp {

// Batched forms used by the chain executor: n same-shape images through device pointer tables.
void launch_pw_f32_batch(cudaStream_t s, const Img &d, const PwProgram &prog, const float *const *in_tab,
                         float *const *out_tab, int n_images, const PwProgram *prog_tab)
{
    const size_t n = d.npix * d.C;
    size_t want = (n / 4 + 96 * 8 - 1) / (96 * 8);
    dim3 grid((unsigned)(want < 1 ? 1 : want), (unsigned)n_images);
    if (prog_tab) {  // one program per image (device memory)
        if (d.C == 1) pw_f32_kernel<1, true><<<grid, 256, 0, s>>>(nullptr, nullptr, n, prog, in_tab, out_tab, prog_tab);
        else if (d.C == 3) pw_f32_kernel<3, true><<<grid, 256, 0, s>>>(nullptr, nullptr, n, prog, in_tab, out_tab, prog_tab);
        else pw_f32_kernel<4, true><<<grid, 256, 0, s>>>(nullptr, nullptr, n, prog, in_tab, out_tab, prog_tab);
    } else if (d.C == 1) pw_f32_kernel<1><<<grid, 256, 0, s>>>(nullptr, nullptr, n, prog, in_tab, out_tab);
    else if (d.C == 3) pw_f32_kernel<3><<<grid, 256, 0, s>>>(nullptr, nullptr, n, prog, in_tab, out_tab);
    else pw_f32_kernel<4><<<grid, 256, 0, s>>>(nullptr, nullptr, n, prog, in_tab, out_tab);
    count_launch();
}

void launch_grey_f32_batch(cudaStream_t s, const Img &d, const PwProgram &pre, const PwProgram &post,
                           const float *const *in_tab, float *const *out_tab, int n_images,
                           const PwProgram *prog_tab)
{
    dim3 grid((unsigned)grey_grid(d), (unsigned)n_images);
    if (prog_tab) {  // [image][2] = (pre, post) per image
        if (d.C == 3) grey_f32_kernel<3, true><<<grid, 256, 0, s>>>(nullptr, nullptr, d.npix, pre, post, in_tab, out_tab, prog_tab);
        else grey_f32_kernel<4, true><<<grid, 256, 0, s>>>(nullptr, nullptr, d.npix, pre, post, in_tab, out_tab, prog_tab);
    } else if (d.C == 3) grey_f32_kernel<3><<<grid, 256, 0, s>>>(nullptr, nullptr, d.npix, pre, post, in_tab, out_tab);
    else grey_f32_kernel<4><<<grid, 256, 0, s>>>(nullptr, nullptr, d.npix, pre, post, in_tab, out_tab);
    count_launch();
}

}  // namespace mp

namespace mp {

}  // namespace mp

namespace mp {

bool fliplr_batch_supported(const Img &d) { return d.W % 128 == 0 && words_per_pixel(d) != 0; }

void launch_fliplr_batch(cudaStream_t s, const Img &d, const void *const *in_tab, void *const *out_tab, int n_images)
{
    const int K = words_per_pixel(d);
    const size_t nu = (size_t)d.H * (d.W / 128);
    size_t blocks = (nu + 7) / 8;
    if (blocks > 65535u * 8u) blocks = 65535u * 8u;
    dim3 grid((unsigned)blocks, (unsigned)n_images);
    const uint4 *const *it = (const uint4 *const *)in_tab;
    uint4 *const *ot = (uint4 *const *)out_tab;
    switch (K) {
        case 1: fliplr_warp_kernel<1><<<grid, 256, 0, s>>>(nullptr, nullptr, d.W, nu, it, ot); break;
        case 2: fliplr_warp_kernel<2><<<grid, 256, 0, s>>>(nullptr, nullptr, d.W, nu, it, ot); break;
        case 3: fliplr_warp_kernel<3><<<grid, 256, 0, s>>>(nullptr, nullptr, d.W, nu, it, ot); break;
        default: fliplr_warp_kernel<4><<<grid, 256, 0, s>>>(nullptr, nullptr, d.W, nu, it, ot); break;
    }
    count_launch();
}

}  // namespace mp

namespace mp {

bool transpose_batch_supported(const Img &d)
{
    const int K = words_per_pixel(d);
    return K != 0 && ((size_t)d.W * K) % 4 == 0 && ((size_t)d.H * K) % 4 == 0;
}

void launch_transpose_batch(cudaStream_t s, const Img &d, const void *const *in_tab, void *const *out_tab,
                            int n_images)
{
    const int K = words_per_pixel(d);
    const uint32_t *const *it = (const uint32_t *const *)in_tab;
    uint32_t *const *ot = (uint32_t *const *)out_tab;
    switch (K) {
        case 1:
            transpose_tma64_kernel<1><<<dim3((d.W + 63) / 64, (d.H + 63) / 64, n_images), 256, 0, s>>>(
                nullptr, nullptr, d.W, d.H, it, ot);
            break;
        case 2:
            transpose_tma64_kernel<2><<<dim3((d.W + 31) / 32, (d.H + 63) / 64, n_images), 256, 0, s>>>(
                nullptr, nullptr, d.W, d.H, it, ot);
            break;
        case 3:
            transpose_tma_kernel<3, 64><<<dim3((d.W + 31) / 32, (d.H + 63) / 64, n_images), 256, 0, s>>>(
                nullptr, nullptr, d.W, d.H, it, ot);
            break;
        default:
            transpose_tma_kernel<4, 64><<<dim3((d.W + 31) / 32, (d.H + 63) / 64, n_images), 256, 0, s>>>(
                nullptr, nullptr, d.W, d.H, it, ot);
            break;
    }
    count_launch();
}

}  // namespace mp
