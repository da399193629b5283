/*
 * mp_image.h -- the eight image operators of the hot path (C ABI).
 *
 * Same symbol names, signature and in-place contract as the reference's
 * src/include/millipyde_image.h:22-44; these are the functions
 * gpuoperation_func_from_name (src/gpuoperation.c:214-257) resolves the strings
 * "rgb2grey", "transpose", "gaussian", "fliplr", "rotate", "brightness",
 * "adjust_gamma", "colorize" to, and the ones every gpuimage method calls
 * (src/gpuimage.c:106, :140, :166, :198, :264, :327, :390, :467).
 *
 * Contract: re-entrant; no CPython calls; select
 * obj->mem_loc themselves; enqueue on obj->stream and return without a host
 * sync; may replace obj->device_data (stream-ordered pool alloc/free) and
 * rewrite ndims/dims/type/nbytes.  `args` is NULL for grey/transpose/fliplr,
 * else the matching *Args struct.  They return a real MPStatus instead of the
 * reference's print-and-exit.
 *
 * Layout dispatch (on obj->type and the channel count):
 *   NPY_UBYTE  H x W x 4   packed RGBA8, reference arithmetic, bit-exact
 *   NPY_UBYTE  H x W x 3   RGB8 (rgb2grey only in the reference; pointwise ops here)
 *   NPY_DOUBLE H x W       reference fp64 greyscale path
 *   NPY_FLOAT  H x W[x1|3|4]  new fp32 path, values in [0, 1], HWC interleaved
 */
#ifndef MP_B200_IMAGE_H
#define MP_B200_IMAGE_H
#include "mp_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

MPStatus mpimg_color_to_greyscale(MPObjData *obj, void *args); /* src/millipyde_image.cpp:532 */
MPStatus mpimg_transpose(MPObjData *obj, void *args);          /* :583 */
MPStatus mpimg_gaussian(MPObjData *obj, void *args);           /* :660, GaussianArgs */
MPStatus mpimg_fliplr(MPObjData *obj, void *args);             /* :678 */
MPStatus mpimg_rotate(MPObjData *obj, void *args);             /* :696, RotateArgs */
MPStatus mpimg_brightness(MPObjData *obj, void *args);         /* :601, BrightnessArgs */
MPStatus mpimg_colorize(MPObjData *obj, void *args);           /* :639, ColorizeArgs */
MPStatus mpimg_adjust_gamma(MPObjData *obj, void *args);       /* :620, GammaArgs */

/*
 * random_* operators (new as C-ABI entry points).  In the reference these exist only as gpuimage
 * methods that draw on the host and call the base method (src/gpuimage.c:206-226, :272-292,
 * :335-355, :398-432, :475-514), so a Pipeline can never run them (its name table,
 * src/gpuoperation.c:214-257, does not know them).  Here they are MPFuncs: each draws with
 * random_double_in_range and calls the base operator, and the chain executor draws per image.
 */
typedef struct { double min, max; } RandomRangeArgs;              /* rotate, gaussian, brightness */
typedef struct { double gamma_min, gamma_max, gain_min, gain_max; } RandomGammaArgs;
typedef struct { double r_min, r_max, g_min, g_max, b_min, b_max; } RandomColorizeArgs;

MPStatus mpimg_random_rotate(MPObjData *obj, void *args);       /* RandomRangeArgs, degrees */
MPStatus mpimg_random_gaussian(MPObjData *obj, void *args);     /* RandomRangeArgs, sigma */
MPStatus mpimg_random_brightness(MPObjData *obj, void *args);   /* RandomRangeArgs, delta */
MPStatus mpimg_random_adjust_gamma(MPObjData *obj, void *args); /* RandomGammaArgs */
MPStatus mpimg_random_colorize(MPObjData *obj, void *args);     /* RandomColorizeArgs */

/* Name -> operator, as gpuoperation_func_from_name (src/gpuoperation.c:214-257) plus the grey
 * spellings (rgb2gray, rgba2grey, rgba2gray) and the random_* names.  *arg_bytes receives the size
 * of the operator's argument block (0 if it takes none).  NULL for an unknown name. */
MPFunc mpimg_func_from_name(const char *name, size_t *arg_bytes);

/*
 * Semantics selector (new).  The reference's float kernels disagree with its own
 * test oracle in two places (findings 2 and 3):
 *   MP_SEMANTICS_ORACLE (default)  float layouts follow the scikit-image calls in
 *       tests/millipyde_tests.py: Gaussian radius int(8*sigma+0.5) with scipy's
 *       float64 weights, bilinear rotate about (W/2-0.5, H/2-0.5), pow in the
 *       image's own precision.
 *   MP_SEMANTICS_REFERENCE  fp64 layouts reproduce the reference kernels bit for
 *       bit: 17 taps from float expf weights + fmax(0), nearest rotate by int
 *       truncation about (W/2, H/2), float powf gamma.
 * RGBA8 always follows the reference arithmetic (it is the only definition and
 * the contract there is bit-exactness); fp32 always follows the oracle.
 */
#define MP_SEMANTICS_ORACLE 0
#define MP_SEMANTICS_REFERENCE 1
void mpimg_set_semantics(int mode);
int mpimg_get_semantics(void);

/* Effective Gaussian support actually evaluated for `sigma` under the oracle
 * semantics (float32 images): taps are dropped from the far end while their total
 * weight stays below 2^-23, one fp32 ulp of 1.0 -- per pass that is less than the
 * rounding of a [0,1] result (sigma = 2: radius 10 of 16, 1.14e-7 dropped).  Returns
 * the radius used; *full is the oracle's nominal radius int(8*sigma+0.5). */
int mpimg_gaussian_effective_radius(double sigma, int *full);

/*
 * Column pass of the fp32 streaming Gaussian (new; a tuning and A/B-test switch, not a
 * semantics switch -- both forms meet the 1e-5 contract and differ by ~1e-6):
 *   MP_GAUSS_COLUMN_MMA (default)  vertical filter on the tensor cores: split-precision
 *       tf32 + fp16-correction mma.sync products, output through TMA bulk stores
 *       (kernels/gaussian_stream_mma.cuh; effective radius <= 11, |sample| < 65504).
 *   MP_GAUSS_COLUMN_FMA  vertical filter on the fp32 FMA pipe
 *       (kernels/gaussian_stream_ws.cuh; every radius bucket, no range limit).
 * The default can also be chosen with the environment variable
 * MILLIPYDE_GAUSS_COLUMN=mma|fma, read on first use.
 */
#define MP_GAUSS_COLUMN_MMA 0
#define MP_GAUSS_COLUMN_FMA 1
void mpimg_set_gauss_column(int mode);
int mpimg_get_gauss_column(void);

/* Work-item plan of one streaming-Gaussian launch (diagnostic, needs no GPU; tests/test_kernel_models.py).
 * `columns` = images x 640-float strips, `radius` = the radius bucket, `sms` = CTAs of the persistent
 * grid.  out = { main_items, tail_cols, chunk_rows, n_chunks }: the first main_items columns are one
 * item each (whole waves of the grid), each of the tail_cols columns behind them is cut into n_chunks
 * row chunks of chunk_rows rows so that the last, partial wave is short. */
void mpimg_gauss_stream_plan(long columns, int height, int radius, int sms, int mma, int out[4]);

/*
 * numpy-ufunc-exact elementwise operators on float32 images (new).  What the
 * extension's __array_ufunc__ dispatches np.add / np.subtract / np.multiply / np.power / np.clip /
 * np.maximum / np.minimum with scalar (or, for multiply, per-channel) operands to, instead of the
 * reference's D2H + CPU path (its __array_ufunc__ is a printing stub, src/gpuarray.c:147-191).
 * No clamp, no special case for alpha: exactly the ufunc.  MPFunc signature, so it chains in a
 * Pipeline like the eight operators.  float32 layouts only (MP_ERROR_UNSUPPORTED_LAYOUT otherwise;
 * MP_EW_MUL with per-channel factors needs <= 3 channels).
 */
#define MP_EW_ADD 4   /* v + a */
#define MP_EW_MUL 5   /* v * a (all channels) or v * (a | b | c) by channel when per_channel != 0 */
#define MP_EW_POW 6   /* powf(v, a) */
#define MP_EW_CLIP 7  /* min(max(v, a), b) */
typedef struct {
    double kind;         /* one of MP_EW_* (a double so the block stays an all-double POD like the reference's *Args) */
    double a, b, c;
    double per_channel;
} ElementwiseArgs;
MPStatus mpimg_elementwise(MPObjData *obj, void *args);

/*
 * Declared value range of fp32 images (new).  The fp32 layout is defined for image data in [0, 1]
 * (the north star's contract, tolerances are stated on it), and under MP_RANGE_UNIT (default) the
 * Gaussian may use the tensor-core column pass, whose fp16 correction operands OVERFLOW for
 * |sample| >= 65504: such samples produce Inf/NaN.  A caller whose float images carry other data
 * (uint16-range, HDR, arbitrary arrays) declares MP_RANGE_ANY: every fp32 Gaussian then runs on the
 * fp32 FMA pipe (no range limit, same 1e-5 contract relative to the data's scale).  Also settable
 * with the environment variable MILLIPYDE_VALUE_RANGE=unit|any, read on first use.
 */
#define MP_RANGE_UNIT 0
#define MP_RANGE_ANY 1
void mpimg_set_value_range(int mode);
int mpimg_get_value_range(void);

#ifdef __cplusplus
}
#endif
#endif /* MP_B200_IMAGE_H */
