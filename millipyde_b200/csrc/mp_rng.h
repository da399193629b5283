// Counter-based random source shared by the host executor and the kernels.
//
// The reference draws every parameter of every image with one getrandom(2) syscall on the submitting
// thread (src/millipyde.c:140-173, called from src/gpuimage.c:206-226, :272-292, :335-355, :398-432,
// :475-514 and src/gpuoperation.c:187-210).  Here a draw is a pure function of
//     (run key, image index, stage index, slot)
// evaluated by Philox-4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11):
// no state, no syscalls, no ordering between threads -- the same stream whichever device or shard an
// image lands on -- and the SAME function on the host (coin flips, sigma -> kernel choice) and in the
// kernel that fills the per-image parameter records on the device (kernels/records.cuh).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MP_HD __host__ __device__ __forceinline__
#else
#define MP_HD inline
#endif

namespace mprng {

struct U4 {
    uint32_t x, y, z, w;
};

MP_HD void mulhilo(uint32_t a, uint32_t b, uint32_t *hi, uint32_t *lo)
{
    const uint64_t p = (uint64_t)a * (uint64_t)b;
    *hi = (uint32_t)(p >> 32);
    *lo = (uint32_t)p;
}

MP_HD U4 philox4x32_10(U4 ctr, uint32_t k0, uint32_t k1)
{
    for (int round = 0; round < 10; ++round) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo(0xD2511F53u, ctr.x, &hi0, &lo0);
        mulhilo(0xCD9E8D57u, ctr.z, &hi1, &lo1);
        ctr = U4{hi1 ^ ctr.y ^ k0, lo1, hi0 ^ ctr.w ^ k1, lo0};
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return ctr;
}

// Slots of a stage: 0..5 = its parameters in declaration order, 7 = its coin (Operation probability).
constexpr uint32_t kCoinSlot = 7;

// Uniform double in [0, 1] with 53 random bits (both ends reachable only by rounding, like the
// reference's buffer / ULONG_MAX).
MP_HD double keyed_u01(uint64_t run_key, uint64_t image, uint32_t stage, uint32_t slot)
{
    const U4 r = philox4x32_10(U4{(uint32_t)image, (uint32_t)(image >> 32), stage, slot}, (uint32_t)run_key,
                               (uint32_t)(run_key >> 32));
    const uint64_t bits = (((uint64_t)r.x << 32) | r.y) >> 11;   // 53 bits
    return (double)bits * (1.0 / 9007199254740992.0);
}

// lo + u * (hi - lo), one rounding (fma) so that host and device agree to the last bit.
MP_HD double keyed_range(uint64_t run_key, uint64_t image, uint32_t stage, uint32_t slot, double lo, double hi)
{
#if defined(__CUDA_ARCH__)
    return fma(keyed_u01(run_key, image, stage, slot), hi - lo, lo);
#else
    return __builtin_fma(keyed_u01(run_key, image, stage, slot), hi - lo, lo);
#endif
}

}  // namespace mprng
