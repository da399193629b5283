// Parameter block of the fused gather segment (kernels/geometry.cuh).
#pragma once
#include "common.cuh"

namespace mpk {

struct IndexMap {  // (y', x') = (ay*y + by*x + cy, ax*y + bx*x + cx)
    int ay, by, cy, ax, bx, cx;
};

// What differs between the images of one batched gather launch when the chain draws random
// parameters: the angle and the pointwise programs (device memory, one record per image).
struct GatherVar {
    RotateParams rp;
    PwProgram pw_pre, pw_post;
};

struct GatherParams {
    const float *const *in_tab;
    float *const *out_tab;
    const float *in;   // single image when the tables are null
    float *out;
    int out_h, out_w;  // output image
    int rot_h, rot_w;  // the image the rotate samples (== its output size); also the bounds for corners
    int src_w;         // row length (pixels) of the source image
    IndexMap post, pre;
    int has_rotate;
    RotateParams rp;
    PwProgram pw_pre, pw_post;
    const GatherVar *var_tab;  // per-image records (kernel template TAB = true), else null
    int pitch;                 // row pitch of the staged box in floats (0: the geometry's default); chosen per angle
};


}  // namespace mpk
