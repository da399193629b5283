// Per-image parameter records filled ON THE DEVICE.
//
// A chain with random_* stages gives every image of a batched launch its own pointwise programs and
// its own rotation.  When the images of a launch share the chain's *shape* (the usual Generator batch:
// same ops, same ranges, only the draws differ) the host uploads ONE template plus the images' stream
// indices, and this kernel evaluates the counter-based generator (mp_rng.h: Philox-4x32-10 keyed by
// run, image index, stage, slot) and writes the records the consumer kernels read -- no host draws, no
// per-image H2D.  The host evaluates the very same function only where it needs a value to pick a
// kernel (a Gaussian's sigma -> radius bucket) or to decide what runs (coin flips).
// Reference: the draws of src/gpuimage.c:206-226, :272-292, :335-355, :398-432, :475-514.
#pragma once
#include "../mp_rng.h"
#include "common.cuh"
#include "gather_params.cuh"

namespace mpk {

// One op of a program template: value_i = lo[i] + u(stage, slot i) * (hi[i] - lo[i]); a fixed op has
// stage == kFixedStage and lo == hi == its value.
constexpr uint32_t kFixedStage = 0xffffffffu;
struct PwTemplateOp {
    int kind;
    uint32_t stage;
    double lo[3], hi[3];
};
struct PwTemplate {
    int n;
    int pad;
    PwTemplateOp ops[kMaxPw];
};

// Gather segment: the rotation's angle (degrees) may be a draw too.
struct GatherTemplate {
    double angle_lo, angle_hi;
    uint32_t angle_stage;   // kFixedStage: angle_lo is the angle
    int width, height;      // of the rotated image (centre of rotation)
    PwTemplate pre, post;
};

__host__ __device__ __forceinline__ PwOp pw_from_template(const PwTemplateOp &t, uint64_t run_key, uint64_t image)
{
    float v[3];
    for (int i = 0; i < 3; ++i)
        v[i] = (float)(t.stage == kFixedStage ? t.lo[i]
                                              : mprng::keyed_range(run_key, image, t.stage, (uint32_t)i, t.lo[i], t.hi[i]));
    return PwOp{t.kind, v[0], v[1], v[2]};
}

// out[image][k] = programs k = 0..per_image-1 of every image, from per_image templates.
__global__ void __launch_bounds__(128)
records_fill_pw_kernel(PwProgram *__restrict__ out, const PwTemplate *__restrict__ tmpl, int per_image,
                       const unsigned long long *__restrict__ image_index, int n_images, unsigned long long run_key)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;   // (image, program, op)
    const int total = n_images * per_image * kMaxPw;
    if (slot >= total) return;
    const int op = slot % kMaxPw, prog = (slot / kMaxPw) % per_image, img = slot / (kMaxPw * per_image);
    const PwTemplate &t = tmpl[prog];
    PwProgram &dst = out[(size_t)img * per_image + prog];
    if (op == 0) dst.n = t.n;
    dst.ops[op] = op < t.n ? pw_from_template(t.ops[op], run_key, image_index[img]) : PwOp{PW_NONE, 0.f, 0.f, 0.f};
}

__global__ void __launch_bounds__(128)
records_fill_gather_kernel(GatherVar *__restrict__ out, const GatherTemplate *__restrict__ tmpl,
                           const unsigned long long *__restrict__ image_index, int n_images, unsigned long long run_key)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;   // (image, 0 = rotation | 1 + op of pre | 1 + kMaxPw + op of post)
    constexpr int kPer = 1 + 2 * kMaxPw;
    if (slot >= n_images * kPer) return;
    const int img = slot / kPer, what = slot % kPer;
    const uint64_t image = image_index[img];
    GatherVar &dst = out[img];
    if (what == 0) {
        const double deg = tmpl->angle_stage == kFixedStage
                               ? tmpl->angle_lo
                               : mprng::keyed_range(run_key, image, tmpl->angle_stage, 0, tmpl->angle_lo, tmpl->angle_hi);
        const double t = deg * (3.14159265358979323846 / 180.0);
        dst.rp.c = cos(t);
        dst.rp.s = sin(t);
        dst.rp.cx = tmpl->width / 2.0 - 0.5;
        dst.rp.cy = tmpl->height / 2.0 - 0.5;
        dst.pw_pre.n = tmpl->pre.n;
        dst.pw_post.n = tmpl->post.n;
        return;
    }
    const bool post = what > kMaxPw;
    const int op = (what - 1) % kMaxPw;
    const PwTemplate &t = post ? tmpl->post : tmpl->pre;
    PwProgram &p = post ? dst.pw_post : dst.pw_pre;
    p.ops[op] = op < t.n ? pw_from_template(t.ops[op], run_key, image) : PwOp{PW_NONE, 0.f, 0.f, 0.f};
}

}  // namespace mpk
