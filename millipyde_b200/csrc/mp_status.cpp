// mperr_str + the host random source.
// Replaces src/millipyde.c:8-138 (message table) and :140-173 (getrandom draws).
#include <limits.h>
#include <sys/random.h>

#include <atomic>
#include <cstdint>

#include "mp_abi.h"
#include "mp_internal_rng.h"
#include "mp_rng.h"

static const char *const kMessages[] = {
#define MP_X(name, msg) msg,
    MP_STATUS_TABLE(MP_X)
#undef MP_X
};

extern "C" const char *mperr_str(MPStatus status)
{
    // MILLIPYDE_SUCCESS has no message in the reference's switch: it falls to default.
    if ((int)status <= 0 || (int)status >= MP_STATUS_COUNT) return "Unknown failure occurred";
    return kMessages[(int)status];
}

// Seeded mode: splitmix64 over an atomic counter -- draw k of a run is a pure
// function of (seed, k), so a replay with the same call order reproduces.
static std::atomic<uint64_t> g_seed{0};
static std::atomic<uint64_t> g_counter{0};

static std::atomic<uint64_t> g_runs{0};   // pipeline runs since the last mprand_seed

extern "C" void mprand_seed(uint64_t seed)
{
    g_seed.store(seed);
    g_counter.store(0);
    g_runs.store(0);
}

namespace mp {

// Key of the k-th pipeline run after mprand_seed(s): a pure function of (s, k), so a seeded program
// replays; unseeded, 8 bytes of getrandom(2) per RUN (the reference: per parameter per image).
static uint64_t run_key_of(uint64_t seed, uint64_t k)
{
    const mprng::U4 r = mprng::philox4x32_10(mprng::U4{(uint32_t)k, (uint32_t)(k >> 32), 0x6d696c6cu, 0x69707964u},
                                             (uint32_t)seed, (uint32_t)(seed >> 32));
    return ((uint64_t)r.x << 32) | r.y;
}

uint64_t next_run_key()
{
    const uint64_t seed = g_seed.load(std::memory_order_relaxed);
    if (seed == 0) {
        unsigned long buf = 0;
        if (getrandom(&buf, sizeof buf, 0) == (ssize_t)sizeof buf) return buf;
        return run_key_of(0x9E3779B97F4A7C15ull, g_runs.fetch_add(1));   // no entropy to be had: still varies per run
    }
    return run_key_of(seed, g_runs.fetch_add(1));
}

uint64_t peek_run_key() { return run_key_of(g_seed.load(), g_runs.load()); }

}  // namespace mp

extern "C" double mprand_keyed_double(uint64_t run_key, uint64_t image, unsigned stage, unsigned slot, double min, double max)
{
    return mprng::keyed_range(run_key, image, stage, slot, min, max);
}

static bool next_u64(uint64_t *out)
{
    uint64_t seed = g_seed.load(std::memory_order_relaxed);
    if (seed == 0) {
        unsigned long buf;
        if (getrandom(&buf, sizeof buf, 0) != (ssize_t)sizeof buf) return false;
        *out = buf;
        return true;
    }
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (g_counter.fetch_add(1) + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    *out = z ^ (z >> 31);
    return true;
}

extern "C" MPStatus random_int_in_range(int min, int max, int *result)
{
    uint64_t r;
    if (!next_u64(&r)) return RAND_ERROR_INSUFFICIENT_BYTES;
    *result = (int)(r % (uint64_t)(max + 1 - min)) + min;
    return MILLIPYDE_SUCCESS;
}

extern "C" MPStatus random_double_in_range(double min, double max, double *result)
{
    uint64_t r;
    if (!next_u64(&r)) return RAND_ERROR_INSUFFICIENT_BYTES;
    double u = (double)r / (double)ULONG_MAX;  // [0, 1], src/millipyde.c:170
    *result = u * (max - min) + min;
    return MILLIPYDE_SUCCESS;
}
