// Device table, streams, stream-ordered pools, target-device state and the
// per-device worker pools.
//
// Replaces src/millipyde_devices.cpp (device table :13-17, init :49-115,
// teardown :122-149, selection :227-323, sync :350-398, P2P probe :424-458)
// and src/millipyde_workers.cpp (FIFO pool).  Re-designed for an 8-GPU NVSwitch
// box: every stream non-blocking; contexts, streams and pool set-up per device on
// first use, peer access + pool access per ordered pair on first hand-off; no
// device resets; allocation through cudaMallocAsync pools.
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <unordered_map>
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

// One entry per visible device.  Only what needs no CUDA context is filled in at
// start-up (SM count, clock, the worker pool); the context, the five streams and
// the pool configuration are created by the first caller that really uses the
// device (`ready`), and peer access is granted pair by pair when a hand-off asks
// for it (`mp::ensure_peer`).  A process that only ever touches one GPU of an
// 8-GPU box therefore owns one context and no peer mappings -- the reference
// creates streams on every device and probes every pair (with device resets) at
// import, src/millipyde_devices.cpp:49-115, :424-458.
struct Device {
    bool valid = false;
    std::once_flag ready_once;
    std::atomic<bool> ready{false};
    cudaStream_t streams[DEVICE_STREAM_COUNT] = {};
    cudaMemPool_t mempool = nullptr;
    work_pool *pool = nullptr;
    int sm_count = 0;
    double perf_metric = 0;  // clockRate x SM count, millipyde_devices.cpp:515-536
    void *flush_buf = nullptr;
    size_t flush_bytes = 0;
};

Device *g_devices = nullptr;   // array of g_count entries (Device holds a once_flag: not movable)
int g_count = 0;
// g_peer[a * n + b]: 0 = not asked yet, 1 = a reaches b's memory (peer access + pool access on),
// 2 = the hardware offers no path
std::vector<char> g_peer;
std::mutex g_peer_mux;
bool g_any_peer = false;
std::atomic<int> g_target{DEVICE_LOC_NO_AFFINITY};
int g_recommended = 0;
std::once_flag g_init_once;
MPStatus g_init_status = MILLIPYDE_SUCCESS;
std::atomic<bool> g_initialized{false};

thread_local char t_last_error[512] = "";

inline bool in_range(int id) { return id >= 0 && id < g_count; }

// Context + streams + pool set-up of one device, once.  Leaves the caller's current device alone.
bool make_ready(int id)
{
    if (!in_range(id) || !g_devices[id].valid) return false;
    Device &d = g_devices[id];
    if (d.ready.load(std::memory_order_acquire)) return true;
    std::call_once(d.ready_once, [&d, id] {
        int prev = -1;
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        bool ok = cudaSetDevice(id) == cudaSuccess;
        for (int s = 0; ok && s < DEVICE_STREAM_COUNT; ++s)
            ok = cudaStreamCreateWithFlags(&d.streams[s], cudaStreamNonBlocking) == cudaSuccess;
        // Keep freed blocks in the pool: steady state never returns memory to the
        // driver, so alloc/free are pure stream-ordered bookkeeping.
        if (ok && cudaDeviceGetDefaultMemPool(&d.mempool, id) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(d.mempool, cudaMemPoolAttrReleaseThreshold, &keep);
        } else {
            d.mempool = nullptr;
        }
        if (!ok) {
            mp::record_cuda_error(cudaGetLastError(), "device set-up", __FILE__, __LINE__);
            d.valid = false;
        }
        if (prev >= 0 && prev != id) cudaSetDevice(prev);
        d.ready.store(ok, std::memory_order_release);
    });
    return d.ready.load(std::memory_order_acquire);
}

MPStatus do_initialize()
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        (void)cudaGetLastError();
        return DEV_ERROR_DEVICE_COUNT;
    }
    g_devices = new Device[count];
    g_count = count;
    g_peer.assign((size_t)count * count, 0);

    double best = -1;
    for (int i = 0; i < count; ++i) {
        Device &d = g_devices[i];
        int khz = 0, sms = 0;   // attribute queries need no context on device i
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, i) != cudaSuccess ||
            cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, i) != cudaSuccess) {
            (void)cudaGetLastError();
            continue;  // DEV_WARN_BAD_DEVICE: skipped by mpdev_get_next_device
        }
        d.sm_count = sms;
        d.perf_metric = (double)khz * sms;
        d.valid = true;
        if (d.perf_metric > best) {
            best = d.perf_metric;
            g_recommended = i;
        }
    }
    if (best < 0) return DEV_ERROR_DEVICE_COUNT;

    // is there any peer path at all?  (a capability query: enables nothing, creates no context)
    for (int a = 0; a < count && !g_any_peer; ++a)
        for (int b = 0; b < count && !g_any_peer; ++b) {
            int ok = 0;
            if (a != b && g_devices[a].valid && g_devices[b].valid &&
                cudaDeviceCanAccessPeer(&ok, a, b) == cudaSuccess && ok)
                g_any_peer = true;
        }
    (void)cudaGetLastError();

    for (int i = 0; i < count; ++i) {
        if (!g_devices[i].valid) continue;
        MPStatus st = mpwrk_create_work_pool(&g_devices[i].pool, THREADS_PER_DEVICE);
        if (st != MILLIPYDE_SUCCESS) return st;
    }
    g_initialized.store(true);
    // the recommended device is the one eager ops land on: have it ready (and current) now
    if (!make_ready(g_recommended)) return DEV_ERROR_DEVICE_COUNT;
    cudaSetDevice(g_recommended);

    // MILLIPYDE_EAGER_PEER=1 (diagnostics): the round-1 behaviour, every ordered pair up front
    const char *eager = getenv("MILLIPYDE_EAGER_PEER");
    if (eager && *eager && *eager != '0')
        for (int a = 0; a < count; ++a)
            for (int b = 0; b < count; ++b)
                if (a != b) mp::ensure_peer(a, b);
    return MILLIPYDE_SUCCESS;
}

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

bool device_ready(int device_id) { return make_ready(device_id); }

// Let device `a` (its kernels and copy engines) reach memory that lives on `b`: peer access a -> b
// and read/write access for `a` on b's pool.  Granted on first request, remembered per ordered pair.
bool ensure_peer(int a, int b)
{
    if (!in_range(a) || !in_range(b)) return false;
    if (a == b) return true;
    std::lock_guard<std::mutex> lk(g_peer_mux);
    char &state = g_peer[(size_t)a * g_count + b];
    if (state) return state == 1;
    state = 2;
    int ok = 0;
    if (cudaDeviceCanAccessPeer(&ok, a, b) != cudaSuccess || !ok) {
        (void)cudaGetLastError();
        return false;
    }
    if (!make_ready(a) || !make_ready(b)) return false;
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    cudaSetDevice(a);
    cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
    (void)cudaGetLastError();
    if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaMemAccessDesc desc = {};
        desc.location.type = cudaMemLocationTypeDevice;
        desc.location.id = a;
        desc.flags = cudaMemAccessFlagsProtReadWrite;
        if (g_devices[b].mempool && cudaMemPoolSetAccess(g_devices[b].mempool, &desc, 1) == cudaSuccess) state = 1;
        else (void)cudaGetLastError();
    }
    if (prev >= 0) cudaSetDevice(prev);
    return state == 1;
}

cudaStream_t device_stream(int device_id, int index)
{
    if (index < 0 || index >= DEVICE_STREAM_COUNT || !make_ready(device_id)) return nullptr;
    return g_devices[device_id].streams[index];
}

cudaStream_t stream_of(const MPObjData *obj)
{
    if (obj->stream) return (cudaStream_t)obj->stream;
    return device_stream(obj->mem_loc, 0);
}

// Stream-ordered block from `pool_device`'s pool.  On out-of-memory the pool's cached blocks may
// be what is in the way (the release threshold is "never"): drain the device, hand the unused part
// of the pool back to the driver and try once more before reporting the failure.
void slab_cache_drop(int only_device);

static void *alloc_from(int pool_device, cudaStream_t stream, size_t nbytes, const char *what, int line)
{
    void *p = nullptr;
    // whole 16-byte vectors: a kernel may read the aligned span that encloses a row (TMA bulk copies of
    // images whose rows are not whole vectors end up to 12 bytes past the last sample)
    nbytes = nbytes ? (nbytes + 15) & ~(size_t)15 : 16;
    if (!make_ready(pool_device) || !g_devices[pool_device].mempool) {
        record_cuda_error(cudaErrorInvalidDevice, what, __FILE__, line);
        return nullptr;
    }
    cudaMemPool_t pool = g_devices[pool_device].mempool;
    cudaError_t e = cudaMallocFromPoolAsync(&p, nbytes, pool, stream);
    if (e == cudaErrorMemoryAllocation) {
        (void)cudaGetLastError();
        slab_cache_drop(pool_device);
        int prev = -1;
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        cudaSetDevice(pool_device);
        cudaDeviceSynchronize();
        cudaMemPoolTrimTo(pool, 0);
        if (prev >= 0 && prev != pool_device) cudaSetDevice(prev);
        (void)cudaGetLastError();
        e = cudaMallocFromPoolAsync(&p, nbytes, pool, stream);
    }
    if (e != cudaSuccess) {
        record_cuda_error(e, what, __FILE__, line);
        return nullptr;
    }
    return p;
}

void *pool_alloc(int device_id, cudaStream_t stream, size_t nbytes)
{
    return alloc_from(device_id, stream, nbytes, "cudaMallocFromPoolAsync", __LINE__);
}

// `stream` may belong to ANOTHER device: access to the pool is granted to it here if it was not yet.
void *pool_alloc_on(int pool_device, cudaStream_t stream, size_t nbytes)
{
    return alloc_from(pool_device, stream, nbytes, "cudaMallocFromPoolAsync(peer)", __LINE__);
}

// ---- slabs: the output buffers of one batched launch in ONE pool allocation -------------------
// A batched launch gives every image a fresh output buffer.  One cudaMallocFromPoolAsync and one
// cudaFreeAsync per image and segment is what a chain of small images costs the host, and those
// driver calls serialise inside a process -- the shards of 8 devices driven by one process ran at a
// third of their single-device rate.  A slab is one allocation cut into equal sub-blocks; a sub-block
// is an ordinary `device_data` pointer and goes back through pool_free like any other, which only
// counts it down; the slab returns to the pool when its last sub-block does -- in the order of that
// free's stream, made to wait (one event each) for the other streams that retired sub-blocks.
namespace {
struct Slab {
    void *base;
    size_t bytes;
    int device;
    int live;                      // sub-blocks not yet retired (under g_slab_mux)
    struct Retired { cudaStream_t stream; int device; } streams[6];
    int n_streams;
};
std::mutex g_slab_mux;
std::unordered_map<void *, Slab *> g_slab_of;   // sub-block -> slab
// Retired slabs are kept here, by device and size, instead of going back to the CUDA pool: the pool
// splits a cached 200 MB block to serve the next 12 MB request, so slabs and single blocks in one
// pool meant a fresh physical allocation (a millisecond) for most slabs and growth until allocations
// failed.  A cached slab carries the event of its release; whoever takes it makes its stream wait.
struct FreeSlab {
    void *base;
    cudaEvent_t released;
};
struct SlabCache {
    std::unordered_map<size_t, std::vector<FreeSlab>> by_bytes;
    size_t bytes = 0;
};
SlabCache g_slab_cache[64];                       // per device, under g_slab_mux
constexpr size_t kSlabCacheBytes = (size_t)64 << 30;   // soft: an allocation that fails drops the cache and retries
const bool g_slabs_on = [] { const char *e = getenv("MILLIPYDE_SLABS"); return !(e && *e == '0'); }();
}  // namespace

bool pool_alloc_many(int device_id, cudaStream_t stream, size_t n, size_t nbytes, void **out)
{
    // small images only: for large ones the driver calls are noise and a slab would pin gigabytes
    // for as long as any one of its images lives
    constexpr size_t kMaxSub = (size_t)32 << 20, kMaxSlab = (size_t)256 << 20;
    if (!g_slabs_on || n < 4 || nbytes == 0 || nbytes > kMaxSub) return false;
    const size_t stride = (nbytes + 255) & ~(size_t)255;
    // Slabs come in three sizes per image size -- 16, 8 or 4 sub-blocks -- and what is left over is
    // allocated singly: the pool caches freed blocks by size (release threshold = never), and slabs
    // of every odd length (group sizes follow the coin flips) fragmented it until allocations failed.
    size_t per = 16;
    while (per > 4 && per * stride > kMaxSlab) per /= 2;
    if (per * stride > kMaxSlab) return false;
    size_t done = 0;
    while (done < n) {
        size_t cnt = per;
        while (cnt > n - done && cnt > 4) cnt /= 2;
        if (cnt > n - done) {   // fewer than 4 left: ordinary blocks
            for (; done < n; ++done) {
                out[done] = alloc_from(device_id, stream, nbytes, "cudaMallocFromPoolAsync", __LINE__);
                if (!out[done]) {
                    for (size_t i = 0; i < done; ++i) pool_free(device_id, stream, out[i]);
                    return false;
                }
            }
            break;
        }
        char *base = nullptr;
        cudaEvent_t released = nullptr;
        if (device_id >= 0 && device_id < 64) {
            std::lock_guard<std::mutex> lk(g_slab_mux);
            auto it = g_slab_cache[device_id].by_bytes.find(cnt * stride);
            if (it != g_slab_cache[device_id].by_bytes.end() && !it->second.empty()) {
                base = (char *)it->second.back().base;
                released = it->second.back().released;
                it->second.pop_back();
                g_slab_cache[device_id].bytes -= cnt * stride;
            }
        }
        if (base) {   // ordered after whatever last used it
            if (released) {
                cudaStreamWaitEvent(stream, released, 0);
                cudaEventDestroy(released);
            }
        } else {
            base = (char *)alloc_from(device_id, stream, cnt * stride, "cudaMallocFromPoolAsync(slab)", __LINE__);
        }
        if (!base) {   // hand back what this call made; the caller falls back to (and reports from) single blocks
            for (size_t i = 0; i < done; ++i) pool_free(device_id, stream, out[i]);
            return false;
        }
        Slab *sl = new Slab{base, cnt * stride, device_id, (int)cnt, {}, 0};
        std::lock_guard<std::mutex> lk(g_slab_mux);
        for (size_t i = 0; i < cnt; ++i) {
            out[done + i] = base + i * stride;
            g_slab_of[out[done + i]] = sl;
        }
        done += cnt;
    }
    return true;
}

void pool_free(int device_id, cudaStream_t stream, void *ptr)
{
    if (!ptr) return;
    Slab *sl = nullptr;
    bool last = false;
    Slab::Retired others[6];
    int n_others = 0;
    if (g_slabs_on) {
        std::lock_guard<std::mutex> lk(g_slab_mux);
        auto it = g_slab_of.find(ptr);
        if (it != g_slab_of.end()) {
            sl = it->second;
            g_slab_of.erase(it);
            bool known = false;
            for (int k = 0; k < sl->n_streams; ++k) known = known || sl->streams[k].stream == stream;
            if (!known && sl->n_streams < 6) sl->streams[sl->n_streams++] = Slab::Retired{stream, device_id};
            else if (!known) sl->n_streams = 7;   // too many to list: drain the device before the release
            last = --sl->live == 0;
            if (last) {
                n_others = sl->n_streams > 6 ? -1 : 0;
                for (int k = 0; n_others >= 0 && k < sl->n_streams; ++k)
                    if (sl->streams[k].stream != stream) others[n_others++] = sl->streams[k];
            }
        }
    }
    if (sl && !last) return;
    if (sl) {
        int prev = -1;
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (n_others < 0) {
            cudaSetDevice(sl->device);
            cudaDeviceSynchronize();
        } else {
            for (int k = 0; k < n_others; ++k) order_after(others[k].device, others[k].stream, stream);
        }
        ptr = sl->base;
        const size_t bytes = sl->bytes;
        const int dev = sl->device;
        delete sl;
        // keep it for the next launch of this size (up to a budget), stamped with its release point
        bool cached = false;
        if (dev >= 0 && dev < 64 && g_initialized.load()) {
            cudaEvent_t ev = nullptr;
            cudaSetDevice(device_id);   // the event belongs to the releasing stream's device
            if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess &&
                cudaEventRecord(ev, stream) == cudaSuccess) {
                std::lock_guard<std::mutex> lk(g_slab_mux);
                if (g_slab_cache[dev].bytes + bytes <= kSlabCacheBytes) {
                    g_slab_cache[dev].by_bytes[bytes].push_back(FreeSlab{ptr, ev});
                    g_slab_cache[dev].bytes += bytes;
                    cached = true;
                }
            }
            if (!cached && ev) cudaEventDestroy(ev);
            (void)cudaGetLastError();
        }
        if (prev >= 0) cudaSetDevice(prev);
        if (cached) return;
    }
    (void)device_id;
    cudaError_t e = cudaFreeAsync(ptr, stream);
    if (e != cudaSuccess) record_cuda_error(e, "cudaFreeAsync", __FILE__, __LINE__);
}

// Give the cached slabs of every device back to their pools (mpdev_trim_pools, teardown, out of memory).
void slab_cache_drop(int only_device)
{
    std::vector<std::pair<int, FreeSlab>> all;
    {
        std::lock_guard<std::mutex> lk(g_slab_mux);
        for (int d = 0; d < 64; ++d) {
            if (only_device >= 0 && d != only_device) continue;
            for (auto &kv : g_slab_cache[d].by_bytes)
                for (FreeSlab &f : kv.second) all.push_back({d, f});
            g_slab_cache[d].by_bytes.clear();
            g_slab_cache[d].bytes = 0;
        }
    }
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    for (auto &df : all) {
        if (cudaSetDevice(df.first) != cudaSuccess) continue;
        cudaStream_t s0 = device_stream(df.first, 0);
        if (df.second.released) {
            cudaStreamWaitEvent(s0, df.second.released, 0);
            cudaEventDestroy(df.second.released);
        }
        cudaFreeAsync(df.second.base, s0);
    }
    (void)cudaGetLastError();
    if (prev >= 0) cudaSetDevice(prev);
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
    for (int i = 0; i < g_count; ++i) {
        Device &d = g_devices[i];
        if (!d.valid) continue;
        if (d.pool) {
            mpwrk_destroy_work_pool(d.pool);
            d.pool = nullptr;
        }
        // Unlike millipyde_devices.cpp:139 no cudaDeviceReset: the context may be
        // shared with other libraries in the process.  Streams and pools die with it.
        if (d.ready.load() && cudaSetDevice(i) == cudaSuccess) cudaDeviceSynchronize();
        d.valid = false;
    }
}

// The reference returns MP_FALSE when P2P *is* supported (inverted predicate,
// millipyde_devices.cpp:156-159); this one answers the question asked.
MPBool mpdev_peer_to_peer_supported(void) { return g_any_peer ? MP_TRUE : MP_FALSE; }

// "Can `device` work on memory that lives on `peer`?"  Answering yes makes it so: the pair's peer
// access and pool access are granted here, on first request (millipyde_devices.cpp:161-167 reads a
// matrix filled at import).
MPBool mpdev_can_use_peer(int device, int peer)
{
    if (!in_range(device) || !in_range(peer)) return MP_FALSE;
    if (!g_devices[device].valid || !g_devices[peer].valid) return MP_FALSE;
    return mp::ensure_peer(device, peer) ? MP_TRUE : MP_FALSE;
}

int mpdev_get_device_count(void) { return g_count; }

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
    if (!g_devices[device_id].ready.load()) return;  // never used: no context to drain (or to create)
    MP_CUDA_WARN(cudaSetDevice(device_id));
    MP_CUDA_WARN(cudaDeviceSynchronize());
}

void mpdev_hard_synchronize_all(void)
{
    // drain every pool first so hand-offs between devices have all been enqueued
    for (int i = 0; i < g_count; ++i)
        if (g_devices[i].valid) mpwrk_work_wait(g_devices[i].pool);
    for (int i = 0; i < g_count; ++i)
        if (g_devices[i].valid) mpdev_hard_synchronize(i);
}

void mpdev_synchronize(void) { MP_CUDA_WARN(cudaDeviceSynchronize()); }

void mpdev_synchronize_all(void)
{
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    for (int i = 0; i < g_count; ++i) {
        if (!g_devices[i].valid || !g_devices[i].ready.load()) continue;
        MP_CUDA_WARN(cudaSetDevice(i));
        MP_CUDA_WARN(cudaDeviceSynchronize());
    }
    if (prev >= 0) cudaSetDevice(prev);
}

// Device.__exit__ calls this after an exception (src/device.c:60-65).  The
// reference resets the whole device, which would free every live gpuimage; here
// the device is drained, the error state cleared and the streams re-created.
void mpdev_reset(int device_id)
{
    if (!mpdev_is_valid_device(device_id) || !g_devices[device_id].ready.load()) return;
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
    for (int i = 0; i < g_count; ++i) {
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
    int n = g_count;
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

MPStatus mpdev_pci_bus_id(int device_id, char *buf, int len)
{
    if (!buf || len < 13 || !in_range(device_id)) return MP_ERROR_INVALID_ARGUMENT;
    MP_CUDA_TRY(cudaDeviceGetPCIBusId(buf, len, device_id));
    return MILLIPYDE_SUCCESS;
}

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

void mpdev_trim_pools(void)
{
    mp::slab_cache_drop(-1);
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    for (int i = 0; i < g_count; ++i) {
        Device &d = g_devices[i];
        if (!d.valid || !d.ready.load() || !d.mempool) continue;
        if (cudaSetDevice(i) != cudaSuccess) continue;
        cudaDeviceSynchronize();
        MP_CUDA_WARN(cudaMemPoolTrimTo(d.mempool, 0));
    }
    (void)cudaGetLastError();
    if (prev >= 0) cudaSetDevice(prev);
}

unsigned long long mpdev_launch_count(void) { return mp::g_launch_count.load(); }

}  // extern "C"
