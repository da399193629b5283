// Separable Gaussian for ANY radius: two passes through a global-memory intermediate.
//
// The last resort behind the streaming kernels (radius buckets up to 13) and the shared-memory tile
// kernels (one tile plus its halo must fit in 200 KB: radius up to ~75 for fp32 RGB, ~70 for fp64):
// the reference accepts every sigma (src/millipyde_image.cpp:660-675), and so must the drop-in.  A
// thread owns one output sample and walks its taps through L1/L2; weights come from device memory
// (the radius is unbounded), accumulation is in double for both sample types.  Zero padding, like the
// reference's kernels (:146-244) and scipy's mode="constant".
#pragma once
#include "common.cuh"

namespace mpk {

// One pass: out[i] = sum_k w[|k|] in[i + k * step] over the taps whose line position stays in [0, n).
//   row pass:    step = C,         position = (i % row_elems) / C, n = W
//   column pass: step = row_elems, position = i / row_elems,       n = H
template <typename T, bool ROWS>
__global__ void __launch_bounds__(256)
gauss_global_pass_kernel(const T *__restrict__ in, T *__restrict__ out, size_t total, int row_elems, int channels,
                         int n_line, const double *__restrict__ w, int radius)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int pos = ROWS ? (int)((i % (size_t)row_elems) / (size_t)channels) : (int)(i / (size_t)row_elems);
        const long step = ROWS ? channels : row_elems;
        const int k_lo = pos < radius ? -pos : -radius;
        const int k_hi = pos + radius >= n_line ? n_line - 1 - pos : radius;
        double acc = 0.0;
        const T *p = in + (long)i + (long)k_lo * step;
        for (int k = k_lo; k <= k_hi; ++k, p += step) acc = fma(__ldg(&w[k < 0 ? -k : k]), (double)__ldg(p), acc);
        out[i] = (T)acc;
    }
}

}  // namespace mpk
