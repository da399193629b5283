/*
 * mp_abi.h -- the operator ABI of the image-augmentation hot path.
 *
 * Layout- and value-compatible with the reference's src/include/millipyde.h:
 * an object file compiled against that header links against libmp_b200.so
 * unchanged.  What each item replaces:
 *
 *   MPObjData        src/include/millipyde.h:16-25  (x86-64: 56 bytes; offsets
 *                    device_data 0, ndims 8, dims 16, type 24, mem_loc 28,
 *                    stream 32, pinned 40, nbytes 48)
 *   MPStatus         src/include/millipyde.h:27-94  (same enumerator order =>
 *                    same numeric values; codes >= MP_STATUS_EXT_BASE are new)
 *   MPFunc           src/include/millipyde.h:96
 *   MPRunnable       src/include/millipyde.h:98-103
 *   *Args            src/include/millipyde.h:105-126 (all-double PODs)
 *   mperr_str, random_*_in_range   src/include/millipyde.h:130-136
 *
 * The status list is kept as one X-macro table so the enum and the message
 * table (csrc/mp_status.cpp) cannot drift apart.
 */
#ifndef MP_B200_ABI_H
#define MP_B200_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MP_TRUE 1
#define MP_FALSE 0
#define MP_UNUSED(x) (void)(x)

/* mem_loc values that are not device ordinals */
#define HOST_LOC (-1)
#define DEVICE_LOC_NO_AFFINITY (-2)

typedef int MPBool;

/* numpy type numbers the ops dispatch on (MPObjData.type).  The reference never
 * reads this field inside an op (only __array__ does, src/gpuarray.c:132); the
 * fp32 layouts are new behind the same ABI. */
#define MP_NPY_UBYTE 2
#define MP_NPY_FLOAT 11
#define MP_NPY_DOUBLE 12

typedef struct {
    void *device_data; /* device buffer, row-major, C-contiguous          */
    int ndims;         /* 2 (H x W) or 3 (H x W x C)                       */
    int *dims;         /* 2*ndims ints: shape, then byte strides           */
    int type;          /* numpy typenum                                    */
    int mem_loc;       /* device ordinal, HOST_LOC, DEVICE_LOC_NO_AFFINITY */
    void *stream;      /* cudaStream_t the next op is enqueued on          */
    MPBool pinned;     /* "pinned to its device" (not page-locked)         */
    size_t nbytes;
} MPObjData;

/* name, message */
#define MP_STATUS_TABLE(X)                                                                         \
    X(MILLIPYDE_SUCCESS, "Success")                                                                \
    X(MOD_ERROR, "Could not import module 'millipyde' due to module creation failure.")            \
    X(MOD_ERROR_CREATE_GPUARRAY_TYPE,                                                              \
      "Could not import module 'millipyde' while creating internal type 'gpuarray'")               \
    X(MOD_ERROR_CREATE_GPUIMAGE_TYPE,                                                              \
      "Could not import module 'millipyde' while creating internal type 'gpuimage'")               \
    X(MOD_ERROR_CREATE_OPERATION_TYPE,                                                             \
      "Could not import module 'millipyde' while creating internal type 'Operation'")              \
    X(MOD_ERROR_CREATE_PIPELINE_TYPE,                                                              \
      "Could not import module 'millipyde' while creating internal type 'Pipeline'")               \
    X(MOD_ERROR_CREATE_DEVICE_TYPE,                                                                \
      "Could not import module 'millipyde' while creating internal type 'Device'")                 \
    X(MOD_ERROR_CREATE_GENERATOR_TYPE,                                                             \
      "Could not import module 'millipyde' while creating internal type 'Generator'")              \
    X(MOD_ERROR_ADD_GPUARRAY,                                                                      \
      "Could not import module 'millipyde' while loading internal type 'gpuarray'")                \
    X(MOD_ERROR_ADD_GPUIMAGE,                                                                      \
      "Could not import module 'millipyde' while loading internal type 'gpuimage'")                \
    X(MOD_ERROR_ADD_OPERATION,                                                                     \
      "Could not import module 'millipyde' while loading internal type 'Operation'")               \
    X(MOD_ERROR_ADD_PIPELINE,                                                                      \
      "Could not import module 'millipyde' while loading internal type 'Pipeline'")                \
    X(MOD_ERROR_ADD_DEVICE,                                                                        \
      "Could not import module 'millipyde' while loading internal type 'Device'")                  \
    X(MOD_ERROR_ADD_GENERATOR,                                                                     \
      "Could not import module 'millipyde' while loading internal type 'Generator'")               \
    X(DEV_ERROR_CURRENT_DEVICE, "GPU runtime failed while querying the current device")            \
    X(DEV_ERROR_DEVICE_COUNT, "GPU runtime failed while querying the device count")                \
    X(DEV_ERROR_DEVICE_PROPERTIES, "GPU runtime failed while querying device properties")          \
    X(DEV_ERROR_PEER_ACCESS_MATRIX_ALLOC,                                                          \
      "Could not allocate internal data-structure 'peer access matrix'")                           \
    X(DEV_ERROR_DEVICE_ARRAY_ALLOC, "Coult not allocate internal data-structure 'device array'")   \
    X(DEV_WARN_NO_PEER_ACCESS,                                                                     \
      "Multiple devices were detected, but peer2peer is not supported on this system")             \
    X(DEV_WARN_BAD_DEVICE, "A device was detected but found to be unusable")                       \
    X(WORK_ERROR_INIT_PTHREAD, "Unable to initialize work pool thread")                            \
    X(WORK_ERROR_INIT_COND, "Unable to initialize work pool condition")                            \
    X(WORK_ERROR_INIT_MUX, "Unable to initialize work pool mutex")                                 \
    X(WORK_ERROR_ALLOC_WORK_POOL, "Unable to allocate memory for work pool")                       \
    X(WORK_ERROR_ALLOC_WORK_NODE, "Unable to allocate memory for worker node")                     \
    X(WORK_ERROR_NULL_WORK_POOL, "The specified work pool was NULL")                               \
    X(TYPE_ERROR_NON_GPUOBJ, "Cannot perform GPU operations on non-GPU-compatible type")           \
    X(RAND_ERROR_INSUFFICIENT_BYTES, "Random number generation failed to supply requrested bytes") \
    X(GPUARRAY_ERROR_CONSTRUCTION_WITHOUT_ARRAY_TYPE,                                              \
      "Constructing gpuarray requires an ndarray or array compatible argument")                    \
    X(GPUARRAY_ERROR_CONSTRUCTION_WITHOUT_NUMERIC_ARRAY,                                           \
      "Constructing gpuarray requires a numeric array type")                                       \
    X(GPUIMAGE_ERROR_CONSTRUCTION_WITHOUT_ARRAY_TYPE,                                              \
      "Construcing gpuimage requires an ndarray or array compatible argument")                     \
    X(GPUIMAGE_ERROR_CONSTRUCTION_WITHOUT_IMAGE_FORMAT,                                            \
      "Construcing gpuimages requires a compatible image format")                                  \
    X(GPUIMAGE_ERROR_CONSTRUCTION_WITHOUT_IMAGE_DIMS,                                              \
      "Construcing gpuimages either a 2 dimensional array image format for single channel "        \
      "images (greyscale), or a 3 dimensional array image format for multi-channel images "        \
      "(rgb/rgba).")                                                                               \
    X(GPUOPERATION_ERROR_CONSTRUCTION_NO_ARGS,                                                     \
      "Contructing Operations requires a runnable/callable argument")                              \
    X(GPUOPERATION_ERROR_INVALID_PROBABILITY,                                                      \
      "Constructing Operation requires a float probability between 0 and 1 (exclusive)")           \
    X(GPUOPERATION_ERROR_CONSTRUCTION_NAMED_ARGS,                                                  \
      "Constructing Operations can only include one named argument designated 'probability'")      \
    X(GPUOPERATION_ERROR_RUN_WITHOUT_STRING_METHOD,                                                \
      "Operations must be constructed with a string method name to run as instance method")        \
    X(GPUOPERATION_ERROR_RUN_UNKNOWN_STRING_METHOD,                                                \
      "Operation's string method name could not be found for the given object")                    \
    X(GPUOPERATION_ERROR_RUN_NO_DEV_RANDOM,                                                        \
      "Unable to use /dev/random for random number generation")                                    \
    X(GPUOPERATION_ERROR_RUN_CANNOT_READ_DEV_RANDOM,                                               \
      "Unable to read from /dev/random for random number generation")                              \
    X(GPUPIPELINE_ERROR_CONSTRUCTION_INVALID_ARGS,                                                 \
      "Constructing Pipeline requires 2 arguments, or 3 arguments for specifying a device")        \
    X(GPUPIPELINE_ERROR_INVALID_DEVICE, "Constructing Pipeline requires an integer device")        \
    X(GPUPIPELINE_ERROR_UNUSABLE_DEVICE,                                                           \
      "Constructing Pipeline requires a device that is useable for GPU operations")                \
    X(GPUPIPELINE_ERROR_CONSTRUCTION_NAMED_ARGS,                                                   \
      "Constructing Pipelines can only include one named argument designated 'device'")            \
    X(GPUPIPELINE_ERROR_NONLIST_INPUTS, "Constructing Pipeline requires a List of inputs")         \
    X(GPUPIPELINE_ERROR_NONLIST_OPERATIONS,                                                        \
      "Constructing Pipeline requires a List of operations")                                       \
    X(GPUPIPELINE_ERROR_NONGPU_INPUT,                                                              \
      "Constructing Pipeline requires all inputs to be GPU compatible")                            \
    X(GPUGENERATOR_ERROR_INVALID_DEVICE, "Constructing Generator requires an integer device")      \
    X(GPUGENERATOR_ERROR_INVALID_MAX,                                                              \
      "Constructing Generator requires an integer number of outputs greater than 0")               \
    X(GPUGENERATOR_ERROR_INVALID_RETURN_TO,                                                        \
      "Constructing Generator requries boolean value for whether to return_to_host")               \
    X(GPUGENERATOR_ERROR_UNUSABLE_DEVICE,                                                          \
      "Constructing Generator requires a device that is useable for GPU operations")               \
    X(GPUGENERATOR_ERROR_CONSTRUCTION_NAMED_ARGS,                                                  \
      "Constructing Generator can only include the named arguments 'device' and/or 'outputs'")     \
    X(GPUGENERATOR_ERROR_INVALID_INPUT,                                                            \
      "Constructing Generator requires an list of inputs or a path to files")                      \
    X(GPUGENERATOR_ERROR_NONLIST_OPERATIONS,                                                       \
      "Constructing Generator requires a list of Operations")                                      \
    /* ---- new codes (not in the reference; its ops print and exit(1) on a runtime error, */     \
    /*      src/include/millipyde_hip_util.h:7-16 -- these return instead) ---------------- */     \
    X(MP_ERROR_CUDA_RUNTIME, "A CUDA runtime call failed (see stderr for the call site)")          \
    X(MP_ERROR_NO_DEVICE, "No usable CUDA device is present")                                      \
    X(MP_ERROR_DEVICE_ALLOC, "Device memory allocation failed")                                    \
    X(MP_ERROR_NULL_DATA, "The GPU object holds no device data")                                   \
    X(MP_ERROR_UNSUPPORTED_LAYOUT,                                                                 \
      "Unsupported image layout: expected uint8 HxWx{3,4}, float64 HxW[x3] or float32 HxW[x{1,3,4}]") \
    X(MP_ERROR_INVALID_ARGUMENT, "Invalid operation argument")                                     \
    X(MP_ERROR_NO_PEER_PATH, "The two devices cannot exchange memory")

typedef enum mp_status_codes {
#define MP_X(name, msg) name,
    MP_STATUS_TABLE(MP_X)
#undef MP_X
        MP_STATUS_COUNT
} MPStatus;

#define MP_STATUS_EXT_BASE MP_ERROR_CUDA_RUNTIME

/* One op of the hot path: mutate `obj` in place (may replace device_data and
 * rewrite ndims/dims/type/nbytes), enqueue on obj->stream, never call CPython. */
typedef MPStatus (*MPFunc)(MPObjData *obj, void *args);

/* One stage of an Operation chain, as Pipeline pre-resolves it
 * (src/gpupipeline.c:152-161).  probability < 0 means "always". */
typedef struct {
    MPFunc func;
    MPObjData *obj_data; /* unused, kept for layout */
    void *args;
    double probability;
} MPRunnable;

typedef struct { double angle; } RotateArgs;                    /* degrees */
typedef struct { double sigma; } GaussianArgs;
typedef struct { double delta; } BrightnessArgs;                /* |delta| < 1 */
typedef struct { double r_mult, g_mult, b_mult; } ColorizeArgs; /* >= 0 */
typedef struct { double gamma, gain; } GammaArgs;

const char *mperr_str(MPStatus status);

/* Uniform draws used by random_* ops and Operation(probability=).  Unseeded
 * they read getrandom(2) like src/millipyde.c:140-173; after mprand_seed(s != 0)
 * they come from a counter-based generator so tests and benchmarks replay. */
MPStatus random_int_in_range(int min, int max, int *result);
MPStatus random_double_in_range(double min, double max, double *result);
void mprand_seed(uint64_t seed); /* new: 0 restores the entropy source */
/* new: the draw a Pipeline / Generator run makes for parameter `slot` (0..5 in declaration order; 7 = the
 * stage's probability coin, on [0, 1]) of stage `stage` of image `image` (its position in the stream) --
 * Philox-4x32-10 keyed by the run (mppipe_last_run_key).  The same function is evaluated on the host and
 * in the kernel that fills per-image parameter records, so a stream can be predicted and replayed. */
double mprand_keyed_double(uint64_t run_key, uint64_t image, unsigned stage, unsigned slot, double min, double max);

#ifdef __cplusplus
}
#endif

#ifdef __cplusplus
static_assert(sizeof(MPObjData) == 56 && offsetof(MPObjData, nbytes) == 48 &&
                  offsetof(MPObjData, stream) == 32 && offsetof(MPObjData, type) == 24,
              "MPObjData must keep the reference layout");
#endif

#endif /* MP_B200_ABI_H */
