/*
 * mp_pipeline.h -- the Operation-chain executor (C ABI): fusion pass, batched
 * launches, device sharding and cross-device hand-off.
 *
 * This is the body of the reference's Pipeline, lifted out of the CPython type:
 *   mppipe_create      <- the MPRunnable[] pre-resolution of PyGPUPipeline_init
 *                         (src/gpupipeline.c:152-161)
 *   mppipe_connect     <- PyGPUPipeline_connect_to (src/gpupipeline.c:186-218)
 *   mppipe_run         <- PyGPUPipeline_run + gpupipeline_run_sequence
 *                         (src/gpupipeline.c:234-312, :352-403)
 *   mppipe_run_host    <- the Generator's clone -> ops -> D2H loop
 *                         (src/gpugenerator.c:203-281), as a pinned, multi-stream
 *                         host->device->host stream
 *
 * What changes behind it (BASELINE.json north star, item 2): the stage loop no
 * longer calls one kernel + one stream sync per stage per image.  Per-image coin
 * flips are drawn first; images with the same layout and the same surviving op
 * list form a group; each group's op list is compiled into segments
 * (index/resample + pointwise -> at most one stencil -> pointwise), and every
 * segment is ONE launch over the whole group (pointer tables), so a chain costs
 * one HBM round trip per segment per image.  Stages whose MPFunc is not one of
 * libmp_b200's eight operators are called one image at a time, as the reference
 * does.
 */
#ifndef MP_B200_PIPELINE_H
#define MP_B200_PIPELINE_H
#include "mp_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mp_pipeline MPPipeline;

/* device_id: a device ordinal, or DEVICE_LOC_NO_AFFINITY ("use the target device if one is set,
 * else spread the inputs over every device", src/gpupipeline.c:245-264).  The stage array and the
 * *Args blocks of libmp_b200's own operators are copied. */
MPPipeline *mppipe_create(const MPRunnable *stages, int num_stages, int device_id);
void mppipe_destroy(MPPipeline *p);

int mppipe_get_device(const MPPipeline *p);
void mppipe_set_device(MPPipeline *p, int device_id);

/* `from`'s results are handed to `to` (NVLink peer copy when the devices differ). */
void mppipe_connect(MPPipeline *from, MPPipeline *to);

/* Run every stage (and every connected pipeline) on every object; returns when all devices
 * involved are idle.  Objects are mutated in place like the eager ops do. */
MPStatus mppipe_run(MPPipeline *p, MPObjData **objs, int n);

/* mppipe_run for views made by mpobj_view_data (each runs on the device its borrowed buffer lives on,
 * one shard per device): the first launch
 * that touches an image reads the borrowed buffer and writes the view's own; a view no stage
 * touched is deep-copied.  On return every view owns its buffer (or holds none, on error). */
MPStatus mppipe_run_views(MPPipeline *p, MPObjData **views, int n);

/* Asynchronous halves of mppipe_run / mppipe_run_views, for callers that drive several pipelines
 * at once (submit B; wait A; submit A; wait B; ...). */
MPStatus mppipe_submit(MPPipeline *p, MPObjData **objs, int n);
MPStatus mppipe_submit_views(MPPipeline *p, MPObjData **views, int n);
MPStatus mppipe_wait(MPPipeline *p);

/* Layout of one result of mppipe_run_host. */
typedef struct {
    int ndims;
    long shape[3];
    int type;      /* numpy typenum */
    size_t nbytes;
    int status;    /* MPStatus of this image */
} MPHostResult;

/* Stream n host images of one layout through the chain: upload (full PCIe rate when host_in is
 * page-locked) -> ops -> download, `depth` images in flight per device on separate streams so
 * copies overlap kernels.  host_out[i] must hold out_capacity bytes. */
MPStatus mppipe_run_host(MPPipeline *p, const void *const *host_in, void *const *host_out,
                         size_t out_capacity, MPHostResult *results, int n, int ndims,
                         const long *shape, int typenum);

/* Random source of a run (new).  Every coin flip and random_* parameter of a run is a pure function of
 * (run key, stream index of the image, stage, slot) -- mprand_keyed_double -- independent of devices,
 * shards and threads.  The stream index of objs[i] is first_index + i: a Generator sets first_index to
 * the number of outputs produced so far, so its stream does not depend on the look-ahead.  The run key is
 * drawn per mppipe_submit / mppipe_run_host (seeded: a function of the seed and the run number).
 * Per-image parameter records are filled on the device when every image of a launch shares the chain's
 * shape (mppipe_set_device_draws(0) evaluates the same draws on the host instead: an A/B switch). */
void mppipe_set_index_base(MPPipeline *p, unsigned long long first_index);
unsigned long long mppipe_last_run_key(const MPPipeline *p);
/* Draw the key now and keep it for every later run of `p`: with mppipe_set_index_base, a Generator's
 * output k is then the same whatever the batch (look-ahead) boundaries are. */
void mppipe_hold_run_key(MPPipeline *p);
void mppipe_set_device_draws(int enabled);
int mppipe_get_device_draws(void);

/* Fusion on (default) / off: off reproduces the reference's one-kernel-per-stage execution with the
 * same kernels, for A/B parity tests. */
void mppipe_set_fusion(int enabled);
int mppipe_get_fusion(void);

/* Kernel launches and fused segments issued by the most recent run of `p` (all devices). */
/* Host-only "explain" (new; needs no GPU): the segments the fusion pass compiles the chain into for
 * an image of numpy type `typenum` with `channels` channels when every stage runs (probabilities are
 * ignored; random_* stages draw their parameters from the seeded source like a real run).  Writes a
 * ';'-separated description into buf, one token per segment = one HBM round trip per image:
 *   pw(op,...)            fused fp32 pointwise program      u8(op,...)   composed RGBA8 byte tables
 *   grey(pre|post)        rgb2grey with the pointwise ops it absorbed
 *   gather(flip,rotate,flip;pre|post)   fliplr / rotate / pointwise ops in one gather pass
 *   gauss(pre|post)       a Gaussian with the pointwise ops before and after it applied inside the
 *                         stencil kernel (on the landed rows / on the finished rows)
 *   op                    an operator run on its own (transpose, gaussian, foreign, ...)
 * Returns the number of segments, or -1 (bad arguments / buffer too small). */
int mppipe_plan(const MPPipeline *p, int typenum, int channels, char *buf, int cap);

unsigned long long mppipe_last_launches(const MPPipeline *p);
int mppipe_last_segments(const MPPipeline *p);

#ifdef __cplusplus
}
#endif
#endif /* MP_B200_PIPELINE_H */
