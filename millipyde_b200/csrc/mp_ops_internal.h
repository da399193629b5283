// Host-side helpers shared by the eager ops (mp_image_ops.cu) and the fused
// chain executor (mp_pipeline.cu).  Internal; not part of the C ABI.
#pragma once
#include <cuda_runtime.h>

#include "kernels/common.cuh"
#include "kernels/gather_params.cuh"
#include "mp_abi.h"
#include "mp_image.h"

namespace mp {

enum Family {
    FAM_RGBA8,      // uint8 H x W x 4, the reference's packed-uint32 path
    FAM_U8_OTHER,   // uint8 with another channel count (rgb2grey only, like the reference)
    FAM_F64,        // float64 H x W, the reference's greyscale path
    FAM_F64_OTHER,  // float64 colour (rgb2grey only)
    FAM_F32,        // float32 H x W [x 1|3|4], the B200 path
};

struct Img {
    int H, W, C;
    int type;      // numpy typenum
    int esize;     // bytes per channel sample
    Family fam;
    size_t npix;
};

bool describe(const MPObjData *o, Img *d);
int words_per_pixel(const Img &d);
int grid_for(int device, size_t work_items, int threads);
int grey_grid(const Img &d);

int oracle_weights(double sigma, double *w, int max_radius);
// Support truncation of the fp32 Gaussians: taps are dropped from the far end while the dropped weight
// (both sides together) stays below one fp32 ulp of 1.0 -- the error that adds per pass on a [0, 1]
// image is below the rounding of the result itself.  sigma = 2: radius 10 of the oracle's 16 (dropped
// mass 1.14e-7; the weights are the oracle's, normalised over all 33 taps, not renormalised).
constexpr double kGaussTailEps = 1.1920928955078125e-07;   // 2^-23
int effective_radius(const double *w, int r, double eps);
void reference_weights(double sigma, double *w_half);

mpk::RotateParams rotate_params(int W, int H, double angle_deg);
mpk::U8Op u8_brightness_op(double delta);

// Both passes of the blur from `in` to `out` (distinct buffers) on stream s.
MPStatus launch_gaussian(int device, cudaStream_t s, const Img &d, const void *in, void *out, double sigma,
                         bool ref_rule);

// Program-level entry points the fusion pass uses (same begin/alloc/launch/retire
// protocol as the eager mpimg_* ops).
MPStatus op_pointwise_f32(MPObjData *obj, const mpk::PwProgram &prog);
MPStatus op_pointwise_rgba8(MPObjData *obj, const mpk::U8Program &prog);
MPStatus op_grey_f32(MPObjData *obj, const mpk::PwProgram &pre, const mpk::PwProgram &post);
// ElementwiseArgs (include/mp_image.h) -> a one-op pointwise program for an image with `channels` channels
MPStatus op_elementwise_program(const ElementwiseArgs *a, int channels, mpk::PwProgram *prog);

// The *_tab record tables (device memory, one entry per image) carry per-image parameters for
// chains with random_* stages; null = the single program / angle passed by value.
void launch_pw_f32_batch(cudaStream_t s, const Img &d, const mpk::PwProgram &prog, const float *const *in_tab,
                         float *const *out_tab, int n_images, const mpk::PwProgram *prog_tab = nullptr);
void launch_grey_f32_batch(cudaStream_t s, const Img &d, const mpk::PwProgram &pre, const mpk::PwProgram &post,
                           const float *const *in_tab, float *const *out_tab, int n_images,
                           const mpk::PwProgram *prog_tab = nullptr);  // [image][2] = pre, post

bool fliplr_batch_supported(const Img &d);
void launch_fliplr_batch(cudaStream_t s, const Img &d, const void *const *in_tab, void *const *out_tab, int n_images);

bool transpose_batch_supported(const Img &d);
void launch_transpose_batch(cudaStream_t s, const Img &d, const void *const *in_tab, void *const *out_tab,
                            int n_images);

// Fused gather segment (kernels/geometry.cuh): n images through the tables in g, or one image.
void launch_gather_f32(cudaStream_t s, int channels, const mpk::GatherParams &g, int n_images);

// fp32 roofline path (kernels/gaussian_stream.cuh)
bool gauss_stream_supported(int W, int C, int radius);
int gauss_stream_bucket(int radius);  // smallest compiled radius >= radius, 0 if none
// n_images <= 64; image i is filtered with gps[i * gps_stride] (stride 0: one sigma for all) and, when
// pw_tab (DEVICE memory) is given, wrapped in the pointwise programs pw_tab[i * pw_stride] (before the
// blur) and pw_tab[i * pw_stride + 1] (after it); pw_stride is 2 (per image) or 0 (one pair for all).
MPStatus launch_gauss_stream_sets(int device, cudaStream_t s, int H, int W, int C, int n_images,
                                  const float *const *in_tab, float *const *out_tab,
                                  const mpk::GaussParams<float> *gps, int gps_stride = 1,
                                  const mpk::PwProgram *pw_tab = nullptr, int pw_stride = 0);
MPStatus launch_gauss_stream(int device, cudaStream_t s, const Img &d, const float *in, float *out,
                             const mpk::GaussParams<float> &gp);
// n_images of one shape in one launch: either a contiguous batch (in/out +
// image_stride floats) or per-image device pointer tables.
MPStatus launch_gauss_stream_batch(int device, cudaStream_t s, int H, int W, int C, int n_images,
                                   const float *in, float *out, size_t image_stride,
                                   const float *const *in_tab, float *const *out_tab,
                                   const mpk::GaussParams<float> &gp);

}  // namespace mp
