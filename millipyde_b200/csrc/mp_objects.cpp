// Device-buffer lifecycle of a GPU object: upload, download, peer move, clone.
//
// Replaces src/millipyde_objects.cpp:19-120.  Differences behind the same
// entry points: pool allocation instead of hipMalloc/hipFree, asynchronous
// copies on the object's stream, event-ordered peer moves over NVLink instead
// of host syncs, no CPython dependency (mpobj_copy_to_host returns malloc()
// memory; the reference returned PyMem_Malloc memory, :43).
#include <cstdlib>
#include <cstring>

#include "mp_internal.h"
#include "mp_objects.h"

namespace {

int default_device()
{
    int dev = mpdev_get_target_device();
    if (dev == DEVICE_LOC_NO_AFFINITY) dev = mpdev_get_recommended_device();
    return dev;
}

size_t elt_size(int typenum)
{
    switch (typenum) {
        case 0: case 1: case 2: return 1;       // bool, int8, uint8
        case 3: case 4: return 2;               // int16, uint16
        case 5: case 6: return 4;               // int32, uint32
        case 7: case 8: case 9: case 10: return 8;  // (u)long, (u)longlong
        case 11: return 4;                      // float32
        case 12: return 8;                      // float64
        case 23: return 2;                      // float16
        default: return 0;
    }
}

// Make `waiter` wait for everything enqueued so far on `signaller` (a stream of
// device `signal_dev`; the event must be created and recorded on that device, the
// waiting stream may live on any device).  Leaves `signal_dev` current.
void chain_streams(int signal_dev, cudaStream_t signaller, cudaStream_t waiter)
{
    if (signaller == waiter) return;
    cudaSetDevice(signal_dev);
    cudaEvent_t ev;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
        (void)cudaGetLastError();
        cudaStreamSynchronize(signaller);
        return;
    }
    MP_CUDA_WARN(cudaEventRecord(ev, signaller));
    MP_CUDA_WARN(cudaStreamWaitEvent(waiter, ev, 0));
    MP_CUDA_WARN(cudaEventDestroy(ev));  // released once the wait has consumed it
}

}  // namespace

namespace mp {
void order_after(int signal_dev, cudaStream_t signaller, cudaStream_t waiter) { chain_streams(signal_dev, signaller, waiter); }
}  // namespace mp

extern "C" {

void mpobj_copy_from_host(MPObjData *obj, void *data, size_t nbytes)
{
    if (mp::ensure_initialized() != MILLIPYDE_SUCCESS) return;
    int dev = default_device();
    MP_CUDA_WARN(cudaSetDevice(dev));
    cudaStream_t s = mp::device_stream(dev, 0);
    if (obj->device_data != NULL) {
        cudaStream_t old = mp::stream_of(obj);
        if (obj->mem_loc >= 0) cudaSetDevice(obj->mem_loc);
        mp::pool_free(obj->mem_loc, old, obj->device_data);
        cudaSetDevice(dev);
        obj->device_data = NULL;
    }
    obj->mem_loc = dev;
    obj->device_data = mp::pool_alloc(dev, s, nbytes);
    obj->nbytes = nbytes;
    if (!obj->device_data) return;
    if (data && nbytes) {
        MP_CUDA_WARN(cudaMemcpyAsync(obj->device_data, data, nbytes, cudaMemcpyHostToDevice, s));
        // The caller's buffer may be pageable numpy memory that goes away right
        // after: complete the copy before returning (as the blocking hipMemcpy did).
        MP_CUDA_WARN(cudaStreamSynchronize(s));
    }
}

MPStatus mpobj_upload_async(MPObjData *obj, const void *src, size_t nbytes)
{
    if (!obj->device_data) return MP_ERROR_NULL_DATA;
    if (nbytes > obj->nbytes) return MP_ERROR_INVALID_ARGUMENT;
    MP_CUDA_TRY(cudaSetDevice(obj->mem_loc));
    MP_CUDA_TRY(cudaMemcpyAsync(obj->device_data, src, nbytes, cudaMemcpyHostToDevice, mp::stream_of(obj)));
    return MILLIPYDE_SUCCESS;
}

MPStatus mpobj_download_async(MPObjData *obj, void *dst, size_t nbytes)
{
    if (!obj->device_data) return MP_ERROR_NULL_DATA;
    if (nbytes > obj->nbytes) return MP_ERROR_INVALID_ARGUMENT;
    MP_CUDA_TRY(cudaSetDevice(obj->mem_loc));
    MP_CUDA_TRY(cudaMemcpyAsync(dst, obj->device_data, nbytes, cudaMemcpyDeviceToHost, mp::stream_of(obj)));
    return MILLIPYDE_SUCCESS;
}

MPStatus mpobj_copy_to_host_into(MPObjData *obj, void *dst, size_t nbytes)
{
    MPStatus st = mpobj_download_async(obj, dst, nbytes);
    if (st != MILLIPYDE_SUCCESS) return st;
    MP_CUDA_TRY(cudaStreamSynchronize(mp::stream_of(obj)));
    return MILLIPYDE_SUCCESS;
}

MPStatus mpobj_synchronize(MPObjData *obj)
{
    if (!obj) return MP_ERROR_INVALID_ARGUMENT;
    if (!obj->device_data || obj->mem_loc < 0) return MILLIPYDE_SUCCESS;
    MP_CUDA_TRY(cudaSetDevice(obj->mem_loc));
    MP_CUDA_TRY(cudaStreamSynchronize(mp::stream_of(obj)));
    return MILLIPYDE_SUCCESS;
}

// mem_loc is left alone: nothing moves, the host gets a copy (millipyde_objects.cpp:39).
void *mpobj_copy_to_host(MPObjData *obj)
{
    void *data = malloc(obj->nbytes ? obj->nbytes : 1);
    if (!data) return NULL;
    if (mpobj_copy_to_host_into(obj, data, obj->nbytes) != MILLIPYDE_SUCCESS) {
        free(data);
        return NULL;
    }
    return data;
}

void mpobj_set_stream(MPObjData *obj, void *new_stream)
{
    cudaStream_t old = mp::stream_of(obj);
    cudaStream_t nxt = (cudaStream_t)new_stream;
    if (obj->device_data && old && nxt && old != nxt) {
        chain_streams(obj->mem_loc, old, nxt);
    }
    obj->stream = new_stream;
}

// Whole-image hand-off over NVLink.  The copy runs on the producer's stream (so
// it is ordered after the kernels that wrote the image) and the consumer-side
// stream waits on an event; the host never blocks.  After the call obj->stream
// is the destination device's stream with the same index.
void mpobj_change_device(MPObjData *obj, int new_dev)
{
    int old_dev = obj->mem_loc;
    if (old_dev == new_dev || obj->device_data == NULL) return;
    if (!mpdev_is_valid_device(new_dev)) return;
    // the copy runs on the source device's stream and writes the destination device's pool memory
    const bool fwd = mpdev_can_use_peer(old_dev, new_dev), back = mpdev_can_use_peer(new_dev, old_dev);
    if (!fwd && !back) {
        // cudaMemcpyPeerAsync still works (staged through the host) -- slower, not fatal.
        static bool warned = false;
        if (!warned) {
            fprintf(stderr, "[millipyde] devices %d and %d have no peer path; staging through host\n",
                    old_dev, new_dev);
            warned = true;
        }
    }
    cudaStream_t src_stream = mp::stream_of(obj);
    int index = 0;
    for (int s = 0; s < DEVICE_STREAM_COUNT; ++s)
        if (mp::device_stream(old_dev, s) == src_stream) index = s;
    cudaStream_t dst_stream = mp::device_stream(new_dev, index);

    MP_CUDA_WARN(cudaSetDevice(new_dev));
    void *dst = mp::pool_alloc(new_dev, dst_stream, obj->nbytes);
    if (!dst) return;
    chain_streams(new_dev, dst_stream, src_stream);  // allocation is ordered on dst_stream
    MP_CUDA_WARN(cudaSetDevice(old_dev));
    MP_CUDA_WARN(cudaMemcpyPeerAsync(dst, new_dev, obj->device_data, old_dev, obj->nbytes, src_stream));
    mp::pool_free(old_dev, src_stream, obj->device_data);
    chain_streams(old_dev, src_stream, dst_stream);  // consumer waits for the copy
    obj->device_data = dst;
    obj->mem_loc = new_dev;
    obj->stream = (void *)dst_stream;
    MP_CUDA_WARN(cudaSetDevice(new_dev));
}

void mpobj_dealloc_device_data(MPObjData *obj)
{
    if (obj == NULL) return;
    if (obj->device_data != NULL && obj->mem_loc >= 0) {
        if (cudaSetDevice(obj->mem_loc) == cudaSuccess)
            mp::pool_free(obj->mem_loc, mp::stream_of(obj), obj->device_data);
        else
            (void)cudaGetLastError();  // after teardown: the context owns the memory
    }
    obj->device_data = NULL;
}

MPObjData *mpobj_clone_data(MPObjData *obj, int device_id, int stream_id)
{
    if (!obj || mp::ensure_initialized() != MILLIPYDE_SUCCESS) return NULL;
    if (!mpdev_is_valid_device(device_id)) device_id = obj->mem_loc;
    MPObjData *c = (MPObjData *)malloc(sizeof(MPObjData));
    if (!c) return NULL;
    c->ndims = obj->ndims;
    c->type = obj->type;
    c->nbytes = obj->nbytes;
    c->pinned = MP_FALSE;
    c->mem_loc = device_id;  // the reference copies obj->mem_loc here (:104) even when cloning across devices
    cudaStream_t cs = mp::device_stream(device_id, stream_id);
    c->stream = (void *)cs;
    c->dims = (int *)malloc(sizeof(int) * 2 * (size_t)c->ndims);
    memcpy(c->dims, obj->dims, sizeof(int) * 2 * (size_t)c->ndims);
    c->device_data = NULL;
    if (obj->device_data) {
        cudaStream_t src_stream = mp::stream_of(obj);
        MP_CUDA_WARN(cudaSetDevice(device_id));
        c->device_data = mp::pool_alloc(device_id, cs, c->nbytes);
        if (c->device_data) {
            chain_streams(obj->mem_loc, src_stream, cs);  // source contents are ready
            cudaSetDevice(device_id);
            if (obj->mem_loc == device_id)
                MP_CUDA_WARN(cudaMemcpyAsync(c->device_data, obj->device_data, c->nbytes,
                                             cudaMemcpyDeviceToDevice, cs));
            else {
                // runs on the clone's stream (device_id) and reads the source device's pool memory
                (void)mpdev_can_use_peer(device_id, obj->mem_loc);
                MP_CUDA_WARN(cudaMemcpyPeerAsync(c->device_data, device_id, obj->device_data,
                                                 obj->mem_loc, c->nbytes, cs));
            }
            // the source may not be freed/overwritten before the copy has read it
            chain_streams(device_id, cs, src_stream);
        }
    }
    return c;
}

MPObjData *mpobj_view_data(MPObjData *obj)
{
    if (!obj || !obj->device_data || mp::ensure_initialized() != MILLIPYDE_SUCCESS) return NULL;
    MPObjData *c = (MPObjData *)malloc(sizeof(MPObjData));
    if (!c) return NULL;
    *c = *obj;
    c->pinned = MP_FALSE;
    int slots = obj->ndims < 3 ? 3 : obj->ndims;
    c->dims = (int *)calloc(2 * (size_t)slots, sizeof(int));
    memcpy(c->dims, obj->dims, sizeof(int) * 2 * (size_t)obj->ndims);
    return c;  // device_data and stream are obj's: the view is ordered after obj's pending work
}

void mpobj_view_rebind(MPObjData *view, MPObjData *src)
{
    if (!view) return;
    mpobj_dealloc_device_data(view);   // stream-ordered: whatever still reads the old result finishes first
    if (!src || !src->device_data) return;
    int *dims = view->dims;            // room for three dimensions (mpobj_view_data)
    *view = *src;
    view->dims = dims;
    view->pinned = MP_FALSE;
    if (src->ndims <= 3) memcpy(dims, src->dims, sizeof(int) * 2 * (size_t)src->ndims);
}

void mpobj_view_rebind_many(MPObjData **views, MPObjData **srcs, int n)
{
    for (int i = 0; views && i < n; ++i) mpobj_view_rebind(views[i], srcs ? srcs[i] : NULL);
}

MPObjData *mpobj_create(const void *host, int ndims, const long *shape, int typenum)
{
    if (mp::ensure_initialized() != MILLIPYDE_SUCCESS) return NULL;
    size_t es = elt_size(typenum);
    if (es == 0 || ndims < 1 || ndims > 8) return NULL;
    MPObjData *o = (MPObjData *)calloc(1, sizeof(MPObjData));
    if (!o) return NULL;
    o->ndims = ndims;
    o->type = typenum;
    o->mem_loc = HOST_LOC;
    o->pinned = MP_FALSE;
    // room for at least 3 dims so rgb2grey (3 -> 2) and nothing else ever needs to grow it
    int slots = ndims < 3 ? 3 : ndims;
    o->dims = (int *)calloc(2 * (size_t)slots, sizeof(int));
    size_t n = es;
    for (int i = ndims - 1; i >= 0; --i) {
        o->dims[i] = (int)shape[i];
        o->dims[ndims + i] = (int)n;
        n *= (size_t)shape[i];
    }
    mpobj_copy_from_host(o, (void *)host, n);
    if (!o->device_data) {
        mpobj_destroy(o);
        return NULL;
    }
    o->stream = mpdev_get_stream(o->mem_loc, 0);
    return o;
}

void mpobj_destroy(MPObjData *obj)
{
    if (!obj) return;
    mpobj_dealloc_device_data(obj);
    free(obj->dims);
    free(obj);
}

void *mphost_alloc_pinned(size_t nbytes)
{
    if (mp::ensure_initialized() != MILLIPYDE_SUCCESS) return NULL;
    void *p = NULL;
    if (cudaHostAlloc(&p, nbytes ? nbytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        (void)cudaGetLastError();
        return NULL;
    }
    return p;
}

void mphost_free_pinned(void *p)
{
    if (p) MP_CUDA_WARN(cudaFreeHost(p));
}

MPStatus mphost_register(void *p, size_t nbytes)
{
    if (mp::ensure_initialized() != MILLIPYDE_SUCCESS) return MP_ERROR_NO_DEVICE;
    MP_CUDA_TRY(cudaHostRegister(p, nbytes, cudaHostRegisterPortable));
    return MILLIPYDE_SUCCESS;
}

MPStatus mphost_unregister(void *p)
{
    MP_CUDA_TRY(cudaHostUnregister(p));
    return MILLIPYDE_SUCCESS;
}

}  // extern "C"
