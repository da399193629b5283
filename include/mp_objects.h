/*
 * mp_objects.h -- device-buffer lifecycle of a GPU object (C ABI).
 *
 * First five entry points: same names and meaning as the reference's
 * src/include/millipyde_objects.h:15-27 (implemented in
 * src/millipyde_objects.cpp:19-120).  Differences behind the ABI: allocations
 * come from a per-device stream-ordered pool (no cudaMalloc/cudaFree per op, no
 * implicit device sync), copies are asynchronous on the object's stream and
 * full-rate when the host buffer is page-locked, peer moves are ordered with
 * events instead of host syncs.
 */
#ifndef MP_B200_OBJECTS_H
#define MP_B200_OBJECTS_H
#include "mp_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* H2D: allocate on the target (else recommended) device and upload.
 * src/millipyde_objects.cpp:19-37 */
void mpobj_copy_from_host(MPObjData *obj, void *data, size_t nbytes);

/* D2H into a fresh malloc() block the caller owns (the reference used
 * PyMem_Malloc, src/millipyde_objects.cpp:41-47; libmp_b200.so is CPython-free). */
void *mpobj_copy_to_host(MPObjData *obj);

/* Peer move onto device_id over NVLink; no-op if already there.
 * src/millipyde_objects.cpp:50-81 */
void mpobj_change_device(MPObjData *obj, int device_id);

/* Return the buffer to its pool. src/millipyde_objects.cpp:84-94 */
void mpobj_dealloc_device_data(MPObjData *obj);

/* Deep copy (possibly cross-device). src/millipyde_objects.cpp:98-120 */
MPObjData *mpobj_clone_data(MPObjData *obj, int device_id, int stream_id);

/* Header copy that BORROWS obj's device buffer (same device): the input of mppipe_run_views and of
 * nothing else -- the executor reads the borrowed buffer in the chain's first launch and gives the
 * view a buffer of its own, which saves the clone's extra pass over the image (the Generator's
 * "clone, then augment", src/gpugenerator.c:236-247).  Never run an eager op on a view or free it
 * while it still borrows; mppipe_run_views leaves no borrowed buffer behind. */
MPObjData *mpobj_view_data(MPObjData *obj);

/* Re-arm a view for another mppipe_run_views / mppipe_submit_views pass over `src`: the buffer the
 * view owns from its previous pass (if any) goes back to the pool in stream order and the view
 * borrows src's buffer and header again.  src == NULL only returns the buffer (the view then holds
 * none and may be destroyed).  Call it on a view that owns its buffer or holds none -- never on one
 * that still borrows.  A Generator epoch over fixed inputs is: rebind, submit, wait, read, repeat. */
void mpobj_view_rebind(MPObjData *view, MPObjData *src);
/* The same for n views at once (srcs == NULL: only return the buffers). */
void mpobj_view_rebind_many(MPObjData **views, MPObjData **srcs, int n);

/* ---- new entry points -------------------------------------------------- */

/* D2H straight into caller memory (e.g. a numpy buffer); waits for completion. */
MPStatus mpobj_copy_to_host_into(MPObjData *obj, void *dst, size_t nbytes);

/* Asynchronous halves used by the staging pipeline (Generator / bench e2e):
 * enqueue the copy on obj->stream and return; the caller syncs the stream. */
MPStatus mpobj_upload_async(MPObjData *obj, const void *src, size_t nbytes);
MPStatus mpobj_download_async(MPObjData *obj, void *dst, size_t nbytes);
/* Wait until everything enqueued for the object (operators, the copies above) has completed. */
MPStatus mpobj_synchronize(MPObjData *obj);

/* Build / destroy a whole MPObjData from C (what src/gpuarray.c:82-114 does
 * inline): shape has ndims entries, strides are derived (C order). `host` may
 * be NULL for an uninitialised device buffer. */
MPObjData *mpobj_create(const void *host, int ndims, const long *shape, int typenum);
void mpobj_destroy(MPObjData *obj);

/* Move the object onto another stream of its device without a host sync: the new
 * stream waits on an event recorded on the old one. */
void mpobj_set_stream(MPObjData *obj, void *new_stream);

/* Page-locked host memory for full-rate PCIe copies. */
void *mphost_alloc_pinned(size_t nbytes);
void mphost_free_pinned(void *p);
MPStatus mphost_register(void *p, size_t nbytes);
MPStatus mphost_unregister(void *p);

/* Last CUDA error text recorded by this thread (empty string if none). */
const char *mp_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* MP_B200_OBJECTS_H */
