// Device table, streams, stream-ordered pools, target-device state and the
// per-device worker pools.
//
// Replaces src/millipyde_devices.cpp (device table :13-17, init :49-115,
// teardown :122-149, selection :227-323, sync :350-398, P2P probe :424-458)
// and src/millipyde_workers.cpp (FIFO pool).  Re-designed for an 8-GPU NVSwitch
// box: every stream non-blocking, peer access + pool access enabled once for all
// ordered pairs, no device resets, allocation through cudaMallocAsync pools.
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "mp_internal.h"
#include "mp_objects.h"

struct work_node {
    MPWorkItem work;
    void *arg;
};

struct work_pool {
    std::mutex mux;
    std::condition_variable work_available;  // workers sleep here
    std::condition_variable idle;            // mpwrk_work_wait sleeps here
    std::deque<work_node *> queue;
    std::vector<std::thread> threads;
    int busy = 0;
    bool running = true;
};

struct mp_event {
    cudaEvent_t ev;
    int device;
};

namespace {

struct Device {
    bool valid = false;
    cudaStream_t streams[DEVICE_STREAM_COUNT] = {};
    work_pool *pool = nullptr;
    int sm_count = 0;
    double perf_metric = 0;  // clockRate x SM count, millipyde_devices.cpp:515-536
    void *flush_buf = nullptr;
    size_t flush_bytes = 0;
};

std::vector<Device> g_devices;
std::vector<char> g_peer;  // g_peer[a * n + b] = a can reach b
bool g_any_peer = false;
std::atomic<int> g_target{DEVICE_LOC_NO_AFFINITY};
int g_recommended = 0;
std::once_flag g_init_once;
MPStatus g_init_status = MILLIPYDE_SUCCESS;
std::atomic<bool> g_initialized{false};

thread_local char t_last_error[512] = "";

MPStatus init_streams(int id)
{
    Device &d = g_devices[id];
    MP_CUDA_TRY(cudaSetDevice(id));
    for (int s = 0; s < DEVICE_STREAM_COUNT; ++s)
        MP_CUDA_TRY(cudaStreamCreateWithFlags(&d.streams[s], cudaStreamNonBlocking));
    return MILLIPYDE_SUCCESS;
}

MPStatus do_initialize()
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        (void)cudaGetLastError();
        return DEV_ERROR_DEVICE_COUNT;
    }
    g_devices.assign(count, Device());
    g_peer.assign((size_t)count * count, 0);

    double best = -1;
    for (int i = 0; i < count; ++i) {
        Device &d = g_devices[i];
        cudaDeviceProp props;
        if (cudaSetDevice(i) != cudaSuccess || cudaGetDeviceProperties(&props, i) != cudaSuccess) {
            (void)cudaGetLastError();
            continue;  // DEV_WARN_BAD_DEVICE: skipped by mpdev_get_next_device
        }
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, i);
        d.sm_count = props.multiProcessorCount;
        d.perf_metric = (double)khz * props.multiProcessorCount;
        if (init_streams(i) != MILLIPYDE_SUCCESS) continue;

        // Keep freed blocks in the pool: steady state never returns memory to the
        // driver, so alloc/free are pure stream-ordered bookkeeping.
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, i) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        d.valid = true;
        if (d.perf_metric > best) {
            best = d.perf_metric;
            g_recommended = i;
        }
    }
    if (best < 0) return DEV_ERROR_DEVICE_COUNT;

    // Peer access for every ordered pair (uniform over NVSwitch).
    for (int a = 0; a < count; ++a) {
        if (!g_devices[a].valid) continue;
        cudaSetDevice(a);
        cudaMemPool_t pool_a = nullptr;
        cudaDeviceGetDefaultMemPool(&pool_a, a);
        for (int b = 0; b < count; ++b) {
            if (a == b || !g_devices[b].valid) continue;
            int ok = 0;
            if (cudaDeviceCanAccessPeer(&ok, a, b) != cudaSuccess || !ok) {
                (void)cudaGetLastError();
                continue;
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                (void)cudaGetLastError();
                continue;
            }
            (void)cudaGetLastError();
            g_peer[(size_t)a * count + b] = 1;
            g_any_peer = true;
            // let device a's kernels and copies touch pool memory that lives on b
            cudaMemPool_t pool_b = nullptr;
            if (cudaDeviceGetDefaultMemPool(&pool_b, b) == cudaSuccess) {
                cudaMemAccessDesc desc = {};
                desc.location.type = cudaMemLocationTypeDevice;
                desc.location.id = a;
                desc.flags = cudaMemAccessFlagsProtReadWrite;
                if (cudaMemPoolSetAccess(pool_b, &desc, 1) != cudaSuccess) (void)cudaGetLastError();
            }
        }
    }

    for (int i = 0; i < count; ++i) {
        if (!g_devices[i].valid) continue;
        MPStatus st = mpwrk_create_work_pool(&g_devices[i].pool, THREADS_PER_DEVICE);
        if (st != MILLIPYDE_SUCCESS) return st;
    }
    cudaSetDevice(g_recommended);
    g_initialized.store(true);
    return MILLIPYDE_SUCCESS;
}

inline bool in_range(int id) { return id >= 0 && id < (int)g_devices.size(); }

}  // namespace

namespace mp {

std::atomic<unsigned long long> g_launch_count{0};

void record_cuda_error(cudaError_t err, const char *expr, const char *file, int line)
{
    snprintf(t_last_error, sizeof t_last_error, "%s: %s (%s) at %s:%d", cudaGetErrorName(err),
             cudaGetErrorString(err), expr, file, line);
    fprintf(stderr, "[millipyde] CUDA error %s\n", t_last_error);
    (void)cudaGetLastError();  // clear the non-sticky part so later calls are not poisoned
}

MPStatus ensure_initialized()
{
    std::call_once(g_init_once, [] { g_init_status = do_initialize(); });
    return g_init_status;
}

cudaStream_t device_stream(int device_id, int index)
{
    if (!in_range(device_id) || index < 0 || index >= DEVICE_STREAM_COUNT) return nullptr;
    return g_devices[device_id].streams[index];
}

cudaStream_t stream_of(const MPObjData *obj)
{
    if (obj->stream) return (cudaStream_t)obj->stream;
    return device_stream(obj->mem_loc, 0);
}

void *pool_alloc(int device_id, cudaStream_t stream, size_t nbytes)
{
    void *p = nullptr;
    if (nbytes == 0) nbytes = 16;
    cudaError_t e = cudaMallocAsync(&p, nbytes, stream);
    if (e != cudaSuccess) {
        record_cuda_error(e, "cudaMallocAsync", __FILE__, __LINE__);
        return nullptr;
    }
    (void)device_id;
    return p;
}

void *pool_alloc_on(int pool_device, cudaStream_t stream, size_t nbytes)
{
    void *p = nullptr;
    if (nbytes == 0) nbytes = 16;
    cudaMemPool_t pool;
    cudaError_t e = cudaDeviceGetDefaultMemPool(&pool, pool_device);
    if (e == cudaSuccess) e = cudaMallocFromPoolAsync(&p, nbytes, pool, stream);
    if (e != cudaSuccess) {
        record_cuda_error(e, "cudaMallocFromPoolAsync", __FILE__, __LINE__);
        return nullptr;
    }
    return p;
}

void pool_free(int device_id, cudaStream_t stream, void *ptr)
{
    (void)device_id;
    if (!ptr) return;
    cudaError_t e = cudaFreeAsync(ptr, stream);
    if (e != cudaSuccess) record_cuda_error(e, "cudaFreeAsync", __FILE__, __LINE__);
}

int sm_count(int device_id) { return in_range(device_id) ? g_devices[device_id].sm_count : 0; }

}  // namespace mp

extern "C" {

const char *mp_last_error(void) { return t_last_error; }

/* ------------------------------------------------------------------ mpdev_* */

MPStatus mpdev_initialize(void) { return mp::ensure_initialized(); }

void mpdev_teardown(void)
{
    if (!g_initialized.exchange(false)) return;
    for (size_t i = 0; i < g_devices.size(); ++i) {
        Device &d = g_devices[i];
        if (!d.valid) continue;
        if (d.pool) {
            mpwrk_destroy_work_pool(d.pool);
            d.pool = nullptr;
        }
        // Unlike millipyde_devices.cpp:139 no cudaDeviceReset: the context may be
        // shared with other libraries in the process.  Streams and pools die with it.
        if (cudaSetDevice((int)i) == cudaSuccess) cudaDeviceSynchronize();
        d.valid = false;
    }
}

// The reference returns MP_FALSE when P2P *is* supported (inverted predicate,
// millipyde_devices.cpp:156-159); this one answers the question asked.
MPBool mpdev_peer_to_peer_supported(void) { return g_any_peer ? MP_TRUE : MP_FALSE; }

MPBool mpdev_can_use_peer(int device, int peer)
{
    if (!in_range(device) || !in_range(peer)) return MP_FALSE;
    if (device == peer) return MP_TRUE;
    return g_peer[(size_t)device * g_devices.size() + peer] ? MP_TRUE : MP_FALSE;
}

int mpdev_get_device_count(void) { return (int)g_devices.size(); }

MPBool mpdev_is_valid_device(int id) { return in_range(id) && g_devices[id].valid ? MP_TRUE : MP_FALSE; }

void *mpdev_get_stream(int device_id, int stream) { return (void *)mp::device_stream(device_id, stream); }

void mpdev_submit_work(int device_id, MPWorkItem work, void *arg)
{
    if (!mpdev_is_valid_device(device_id)) return;
    mpwrk_work_queue_push(g_devices[device_id].pool, work, arg);
}

void mpdev_hard_synchronize(int device_id)
{
    if (!mpdev_is_valid_device(device_id)) return;
    mpwrk_work_wait(g_devices[device_id].pool);
    MP_CUDA_WARN(cudaSetDevice(device_id));
    MP_CUDA_WARN(cudaDeviceSynchronize());
}

void mpdev_hard_synchronize_all(void)
{
    // drain every pool first so hand-offs between devices have all been enqueued
    for (size_t i = 0; i < g_devices.size(); ++i)
        if (g_devices[i].valid) mpwrk_work_wait(g_devices[i].pool);
    for (size_t i = 0; i < g_devices.size(); ++i)
        if (g_devices[i].valid) mpdev_hard_synchronize((int)i);
}

void mpdev_synchronize(void) { MP_CUDA_WARN(cudaDeviceSynchronize()); }

void mpdev_synchronize_all(void)
{
    for (size_t i = 0; i < g_devices.size(); ++i) {
        if (!g_devices[i].valid) continue;
        MP_CUDA_WARN(cudaSetDevice((int)i));
        MP_CUDA_WARN(cudaDeviceSynchronize());
    }
}

// Device.__exit__ calls this after an exception (src/device.c:60-65).  The
// reference resets the whole device, which would free every live gpuimage; here
// the device is drained, the error state cleared and the streams re-created.
void mpdev_reset(int device_id)
{
    if (!mpdev_is_valid_device(device_id)) return;
    MP_CUDA_WARN(cudaSetDevice(device_id));
    cudaDeviceSynchronize();
    (void)cudaGetLastError();
}

void mpdev_set_device(int device_id) { MP_CUDA_WARN(cudaSetDevice(device_id)); }

void mpdev_stream_synchronize(int device_id, int stream_id)
{
    cudaStream_t s = mp::device_stream(device_id, stream_id);
    if (s) MP_CUDA_WARN(cudaStreamSynchronize(s));
}

int mpdev_get_target_device(void) { return g_target.load(); }

void mpdev_set_target_device(int device_id)
{
    // DEVICE_LOC_NO_AFFINITY restores "no target" (Device.__exit__ passes the saved value back)
    if (device_id == DEVICE_LOC_NO_AFFINITY || mpdev_is_valid_device(device_id)) g_target.store(device_id);
}

int mpdev_get_recommended_device(void) { return g_recommended; }

int mpdev_get_alternative_device(int device_id)
{
    int alt = DEVICE_LOC_NO_AFFINITY;
    double best = 0;
    for (int i = 0; i < (int)g_devices.size(); ++i) {
        if (i == device_id || !g_devices[i].valid) continue;
        if (g_devices[i].perf_metric > best) {
            best = g_devices[i].perf_metric;
            alt = i;
        }
    }
    return alt;
}

int mpdev_get_next_device(int device_id)
{
    int n = (int)g_devices.size();
    for (int i = 0; i < n; ++i) {
        int id = (i + 1 + device_id) % n;
        if (g_devices[id].valid) return id;
    }
    return device_id;
}

/* ------------------------------------------------------------------ mpwrk_* */

MPStatus mpwrk_create_work_node(MPWorkNode **result, MPWorkItem work, void *arg)
{
    work_node *n = new (std::nothrow) work_node{work, arg};
    if (!n) return WORK_ERROR_ALLOC_WORK_NODE;
    *result = n;
    return MILLIPYDE_SUCCESS;
}

void mpwrk_destroy_work_node(MPWorkNode *node) { delete node; }

// Caller holds pool->mux (same rule as millipyde_workers.cpp:57-83).
MPWorkNode *mpwrk_work_queue_pop(MPDeviceWorkPool *pool)
{
    if (!pool || pool->queue.empty()) return nullptr;
    work_node *n = pool->queue.front();
    pool->queue.pop_front();
    return n;
}

MPStatus mpwrk_work_queue_push(MPDeviceWorkPool *pool, MPWorkItem work, void *arg)
{
    if (!pool) return WORK_ERROR_NULL_WORK_POOL;
    work_node *n = nullptr;
    MPStatus st = mpwrk_create_work_node(&n, work, arg);
    if (st != MILLIPYDE_SUCCESS) return st;
    {
        std::lock_guard<std::mutex> lk(pool->mux);
        pool->queue.push_back(n);
    }
    pool->work_available.notify_one();
    return MILLIPYDE_SUCCESS;
}

void mpwrk_work_wait(MPDeviceWorkPool *pool)
{
    if (!pool) return;
    std::unique_lock<std::mutex> lk(pool->mux);
    pool->idle.wait(lk, [pool] { return pool->queue.empty() && pool->busy == 0; });
}

void *mpwrk_process_work(void *arg)
{
    work_pool *pool = (work_pool *)arg;
    for (;;) {
        work_node *n;
        {
            std::unique_lock<std::mutex> lk(pool->mux);
            pool->work_available.wait(lk, [pool] { return !pool->running || !pool->queue.empty(); });
            if (pool->queue.empty()) break;  // stopping and drained
            n = mpwrk_work_queue_pop(pool);
            ++pool->busy;
        }
        n->work(n->arg);
        mpwrk_destroy_work_node(n);
        {
            std::lock_guard<std::mutex> lk(pool->mux);
            --pool->busy;
            if (pool->queue.empty() && pool->busy == 0) pool->idle.notify_all();
        }
    }
    return nullptr;
}

MPStatus mpwrk_create_work_pool(MPDeviceWorkPool **result, int num_threads)
{
    work_pool *pool = new (std::nothrow) work_pool();
    if (!pool) return WORK_ERROR_ALLOC_WORK_POOL;
    try {
        for (int i = 0; i < num_threads; ++i) pool->threads.emplace_back(mpwrk_process_work, pool);
    } catch (...) {
        mpwrk_destroy_work_pool(pool);
        return WORK_ERROR_INIT_PTHREAD;
    }
    *result = pool;
    return MILLIPYDE_SUCCESS;
}

MPStatus mpwrk_destroy_work_pool(MPDeviceWorkPool *pool)
{
    if (!pool) return WORK_ERROR_NULL_WORK_POOL;
    {
        std::lock_guard<std::mutex> lk(pool->mux);
        pool->running = false;
    }
    pool->work_available.notify_all();
    for (auto &t : pool->threads)
        if (t.joinable()) t.join();
    for (work_node *n : pool->queue) delete n;
    delete pool;
    return MILLIPYDE_SUCCESS;
}

/* ------------------------------------------------------- timing / queries */

MPEvent *mpdev_event_create(int device_id)
{
    if (mp::ensure_initialized() != MILLIPYDE_SUCCESS) return nullptr;
    if (cudaSetDevice(device_id) != cudaSuccess) return nullptr;
    mp_event *e = new mp_event{nullptr, device_id};
    if (cudaEventCreate(&e->ev) != cudaSuccess) {
        delete e;
        return nullptr;
    }
    return e;
}

void mpdev_event_destroy(MPEvent *ev)
{
    if (!ev) return;
    cudaEventDestroy(ev->ev);
    delete ev;
}

void mpdev_event_record(MPEvent *ev, void *stream)
{
    if (!ev) return;
    MP_CUDA_WARN(cudaSetDevice(ev->device));
    MP_CUDA_WARN(cudaEventRecord(ev->ev, (cudaStream_t)stream));
}

float mpdev_event_elapsed_ms(MPEvent *start, MPEvent *stop)
{
    float ms = -1.f;
    if (!start || !stop) return ms;
    MP_CUDA_WARN(cudaEventSynchronize(stop->ev));
    MP_CUDA_WARN(cudaEventElapsedTime(&ms, start->ev, stop->ev));
    return ms;
}

MPStatus mpdev_mem_info(int device_id, size_t *free_bytes, size_t *total_bytes)
{
    MP_CUDA_TRY(cudaSetDevice(device_id));
    MP_CUDA_TRY(cudaMemGetInfo(free_bytes, total_bytes));
    return MILLIPYDE_SUCCESS;
}

int mpdev_sm_count(int device_id) { return mp::sm_count(device_id); }

void mpdev_flush_l2(int device_id, void *stream)
{
    if (!mpdev_is_valid_device(device_id)) return;
    Device &d = g_devices[device_id];
    MP_CUDA_WARN(cudaSetDevice(device_id));
    if (!d.flush_buf) {
        d.flush_bytes = (size_t)256 << 20;  // 2x the 126 MB L2
        if (cudaMalloc(&d.flush_buf, d.flush_bytes) != cudaSuccess) {
            d.flush_buf = nullptr;
            return;
        }
    }
    MP_CUDA_WARN(cudaMemsetAsync(d.flush_buf, 0, d.flush_bytes, (cudaStream_t)stream));
}

unsigned long long mpdev_launch_count(void) { return mp::g_launch_count.load(); }

}  // extern "C"
