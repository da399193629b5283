// Internal glue shared by the translation units of libmp_b200.so.  Not part of
// the C ABI (see include/*.h for that).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>

#include "mp_abi.h"
#include "mp_devices.h"

namespace mp {

// Record a failed CUDA call for mp_last_error() and stderr; never exits (the
// reference's HIP_CHECK prints and exit(1)s, src/include/millipyde_hip_util.h:7-16).
void record_cuda_error(cudaError_t err, const char *expr, const char *file, int line);

// Lazily runs mpdev_initialize() once; returns its status.
MPStatus ensure_initialized();

// Streams of a device; the first call for a device creates its context, its five streams and
// configures its pool (devices nobody uses cost nothing).  nullptr for an invalid device.
cudaStream_t device_stream(int device_id, int index);
bool device_ready(int device_id);

// Grant device `a` access to memory living on `b` (peer access a -> b plus read/write access on b's
// pool), on first request per ordered pair.  false: no hardware path.
bool ensure_peer(int a, int b);

// Per-device stream-ordered pool (cudaMallocAsync on the device's default pool
// with the release threshold lifted, so steady state never calls the OS).
void *pool_alloc(int device_id, cudaStream_t stream, size_t nbytes);
void pool_free(int device_id, cudaStream_t stream, void *ptr);
// n equal blocks for the outputs of one batched launch, cut from one pool allocation (a "slab":
// mp_devices.cpp).  Each out[i] is an ordinary buffer pointer and is released with pool_free.  false
// (and nothing allocated) when the request is not worth a slab or cannot be served: allocate singly.
bool pool_alloc_many(int device_id, cudaStream_t stream, size_t n, size_t nbytes, void **out);
// Block from `pool_device`'s pool, allocated in the order of `stream`, which may belong to ANOTHER
// device that has been granted access to the pool (mp::ensure_peer(stream's device, pool_device)):
// a producer kernel can then write its result straight into the consumer device's memory over NVLink.
void *pool_alloc_on(int pool_device, cudaStream_t stream, size_t nbytes);

// The stream an op should use for `obj`: obj->stream, or the device's stream 0
// when a foreign caller left it NULL.
cudaStream_t stream_of(const MPObjData *obj);

// Make `waiter` wait for everything enqueued so far on `signaller` (a stream of device signal_dev):
// one event, no host sync.
void order_after(int signal_dev, cudaStream_t signaller, cudaStream_t waiter);

extern std::atomic<unsigned long long> g_launch_count;
inline void count_launch(unsigned n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

int sm_count(int device_id);

}  // namespace mp

#define MP_CUDA_TRY(expr)                                           \
    do {                                                            \
        cudaError_t mp_e_ = (expr);                                 \
        if (mp_e_ != cudaSuccess) {                                 \
            mp::record_cuda_error(mp_e_, #expr, __FILE__, __LINE__); \
            return MP_ERROR_CUDA_RUNTIME;                           \
        }                                                           \
    } while (0)

// For void functions of the reference ABI (mpobj_*, mpdev_*): record and go on.
#define MP_CUDA_WARN(expr)                                          \
    do {                                                            \
        cudaError_t mp_e_ = (expr);                                 \
        if (mp_e_ != cudaSuccess)                                   \
            mp::record_cuda_error(mp_e_, #expr, __FILE__, __LINE__); \
    } while (0)
