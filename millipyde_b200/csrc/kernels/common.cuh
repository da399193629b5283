// Shared device helpers: streaming 128-bit accesses and the pointwise op
// "program" that both the eager ops and the fused chain kernels execute.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpk {

constexpr int kMaxPw = 8;  // pointwise ops one fused segment can carry

enum PwKind : int {
    PW_NONE = 0,
    PW_BRIGHTNESS = 1,  // a = delta                     src/millipyde_image.cpp:384-398
    PW_GAMMA = 2,       // a = gamma, b = gain           :438-454
    PW_COLORIZE = 3,    // a, b, c = r, g, b multipliers :494-524 (float analogue)
    // numpy-ufunc-exact forms (no clamp, alpha treated like any channel): what __array_ufunc__
    // dispatches np.add / np.multiply / np.power / np.clip on a float32 gpuimage to
    // (reference: the printing stub src/gpuarray.c:147-191 -- there these run on the host)
    PW_EW_ADD = 4,      // v + a
    PW_EW_MUL = 5,      // v * (a | b | c by channel; a for single-channel images)
    PW_EW_POW = 6,      // powf(v, a)
    PW_EW_CLIP = 7,     // min(max(v, a), b)
};

struct PwOp {
    int kind;
    float a, b, c;
};

struct PwProgram {
    int n;
    PwOp ops[kMaxPw];
};

// Streaming (evict-first) 128-bit global accesses: every byte of an image is
// touched once per op, so nothing should linger in L1/L2 ahead of the halo rows
// the stencil kernels do want cached.
__device__ __forceinline__ float4 ld_stream(const float4 *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float4 *p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ uint4 ld_stream(const uint4 *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(uint4 *p, uint4 v) { __stcs(p, v); }
__device__ __forceinline__ double2 ld_stream(const double2 *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(double2 *p, double2 v) { __stcs(p, v); }

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

// x^g for image samples (x in [0, 1], g > 0) as exp2(g * log2 x): two MUFU ops instead of the
// ~40-instruction powf, which made gamma compute-bound (34 % of HBM).  Absolute error on [0, 1]:
// |d(x^g)| = x^g ln2 |d(g log2 x)| and x^g g |log2 x| <= 1/(e ln 2), so the 2^-22 relative error
// of lg2.approx costs < 2e-7 for every g -- far inside the 1e-5 contract (tests check it for
// g in {0.5, 1.5, 2, 2.2}).  x = 0 gives exp2(-inf) = 0; x < 0 gives NaN like powf.
// The .ftz forms are single MUFU ops (no denormal pre/post-scaling): a denormal x counts as 0 and a
// denormal result is 0, an absolute error below 1.2e-38.
__device__ __forceinline__ float pow01(float x, float g)
{
    float l, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(g * l));
    return r;
}

// One pointwise op on one fp32 sample of channel `ch` of a C-channel image.
// Alpha (ch == 3) is never touched, as in the reference's RGBA kernels.
template <int C>
__device__ __forceinline__ float pw_apply_one(const PwOp &op, float v, int ch)
{
    if (C == 4 && ch == 3) return v;
    switch (op.kind) {
        case PW_BRIGHTNESS: return clamp01(v + op.a);
        case PW_GAMMA: return clamp01(op.b * pow01(v, op.a));
        case PW_COLORIZE:
            if (C >= 3) return fminf(1.f, v * (ch == 0 ? op.a : (ch == 1 ? op.b : op.c)));
            return v;
        default: return v;
    }
}
// accurate powf, out of line: inlined it is ~150 instructions per unrolled sample in every kernel that
// can run a pointwise program, for an op only np.power uses
static __device__ __noinline__ float pw_powf(float x, float y) { return powf(x, y); }

// the ufunc-exact kinds, shared by the scalar and the tile form (every channel, alpha included)
__device__ __forceinline__ float pw_ew_one(const PwOp &op, float v, int ch)
{
    switch (op.kind) {
        case PW_EW_ADD: return v + op.a;
        case PW_EW_MUL: return v * (ch == 0 ? op.a : (ch == 1 ? op.b : op.c));
        case PW_EW_POW: return pw_powf(v, op.a);
        default: return fminf(fmaxf(v, op.a), op.b);
    }
}

template <int C>
__device__ __forceinline__ float pw_apply(const PwProgram &prog, float v, int ch)
{
    for (int i = 0; i < prog.n; ++i)
        v = prog.ops[i].kind >= PW_EW_ADD ? pw_ew_one(prog.ops[i], v, C == 1 ? 0 : ch) : pw_apply_one<C>(prog.ops[i], v, ch);
    return v;
}

// The same program applied to a register tile of N samples whose channels are ch0, ch0+1, ...
// (mod C): the op switch is taken once per op, not once per sample.
// channel of element k of a run that starts at channel ch0 (0 <= ch0 < C): no integer division
template <int C>
__device__ __forceinline__ int pw_channel(int ch0, int k)
{
    if (C == 1) return 0;
    if (C == 4) return (ch0 + k) & 3;
    const int t = ch0 + (k % 3);
    return t >= 3 ? t - 3 : t;
}

template <int C, int N>
__device__ __forceinline__ void pw_apply_op_tile_impl(const PwOp &op, float (&r)[N], int ch0)
{
    {
        switch (op.kind) {
            case PW_BRIGHTNESS:
#pragma unroll
                for (int k = 0; k < N; ++k)
                    if (!(C == 4 && ((ch0 + k) & 3) == 3)) r[k] = clamp01(r[k] + op.a);
                break;
            case PW_GAMMA:
#pragma unroll
                for (int k = 0; k < N; ++k)
                    if (!(C == 4 && ((ch0 + k) & 3) == 3)) r[k] = clamp01(op.b * pow01(r[k], op.a));
                break;
            case PW_COLORIZE:
                if (C == 3) {   // the three factors in the run's channel order, selected once
                    const float m[3] = {ch0 == 0 ? op.a : (ch0 == 1 ? op.b : op.c), ch0 == 0 ? op.b : (ch0 == 1 ? op.c : op.a),
                                        ch0 == 0 ? op.c : (ch0 == 1 ? op.a : op.b)};
#pragma unroll
                    for (int k = 0; k < N; ++k) r[k] = fminf(1.f, r[k] * m[k % 3]);
                } else if (C == 4) {
#pragma unroll
                    for (int k = 0; k < N; ++k) {
                        const int ch = pw_channel<C>(ch0, k);
                        if (ch < 3) r[k] = fminf(1.f, r[k] * (ch == 0 ? op.a : (ch == 1 ? op.b : op.c)));
                    }
                }
                break;
            case PW_EW_ADD:
#pragma unroll
                for (int k = 0; k < N; ++k) r[k] = r[k] + op.a;
                break;
            case PW_EW_MUL:
                if (C == 3) {
                    const float m[3] = {ch0 == 0 ? op.a : (ch0 == 1 ? op.b : op.c), ch0 == 0 ? op.b : (ch0 == 1 ? op.c : op.a),
                                        ch0 == 0 ? op.c : (ch0 == 1 ? op.a : op.b)};
#pragma unroll
                    for (int k = 0; k < N; ++k) r[k] = r[k] * m[k % 3];
                } else {   // one factor for every channel (per-channel factors need exactly three channels)
#pragma unroll
                    for (int k = 0; k < N; ++k) r[k] = r[k] * op.a;
                }
                break;
            case PW_EW_CLIP:
#pragma unroll
                for (int k = 0; k < N; ++k) r[k] = fminf(fmaxf(r[k], op.a), op.b);
                break;
            case PW_EW_POW:   // out of line (see pw_powf): the only case with a call in it
#pragma unroll
                for (int k = 0; k < N; ++k) r[k] = pw_powf(r[k], op.a);
                break;
            default: break;
        }
    }
}
template <int C, int N>
__device__ __forceinline__ void pw_apply_tile(const PwProgram &prog, float (&r)[N], int ch0)
{
    for (int i = 0; i < prog.n; ++i) pw_apply_op_tile_impl<C, N>(prog.ops[i], r, ch0);
}

// A program staged in shared memory for kernels that apply it many times (the streaming Gaussian's
// fused pre/post ops): ops on 16-byte boundaries so that one op is one LDS.128 broadcast.
struct alignas(16) PwSmem {
    int n;
    int pad[3];
    int4 ops[kMaxPw];
};
// one warp copies `src` (device memory; null = empty program) into its own PwSmem
__device__ __forceinline__ void pw_smem_load(PwSmem *dst, const PwProgram *__restrict__ src, int lane)
{
    if (lane == 0) dst->n = src ? __ldg(&src->n) : 0;
    if (src) reinterpret_cast<int *>(dst->ops)[lane] = __ldg(reinterpret_cast<const int *>(src->ops) + lane);  // 8 ops x 4 words
    __syncwarp();
}
__device__ __forceinline__ PwOp pw_smem_op(const PwSmem &p, int i)
{
    const int4 raw = p.ops[i];
    return PwOp{raw.x, __int_as_float(raw.y), __int_as_float(raw.z), __int_as_float(raw.w)};
}
// one op on a register tile (the body of pw_apply_tile for a single op)
template <int C, int N>
__device__ __forceinline__ void pw_apply_op_tile(const PwOp &op, float (&r)[N], int ch0)
{
    pw_apply_op_tile_impl<C, N>(op, r, ch0);
}

// Luma of skimage.color.rgb2gray / src/millipyde_image.cpp:64, fp32 flavour.
__device__ __forceinline__ float luma_f32(float r, float g, float b)
{
    return fminf(1.f, fmaf(0.0721f, b, fmaf(0.7154f, g, 0.2125f * r)));
}


// ---- parameter blocks shared between host launch code and the kernels -------

// fp64 pointwise op (reference greyscale layout)
struct PwOp64 {
    int kind;
    double a, b;
};

// One byte -> byte op of the packed-RGBA8 path.
struct U8Op {
    int kind;
    int d8;          // brightness: (char)(delta * 255), computed on the host (:613)
    double a, b, c;  // gamma: a = gamma, b = gain; colorize: r, g, b multipliers
};

struct U8Program {
    int n;
    U8Op ops[kMaxPw];
};

// Bilinear rotate: cos/sin of the angle and the centre, all fp64.
struct RotateParams {
    double c, s;    // cos, sin of the angle
    double cx, cy;  // centre
};

constexpr int kGaussMaxRadius = 127;

// Symmetric 1-D kernel: w[d] = weight at distance d.
template <typename T>
struct GaussParams {
    int radius;
    T w[kGaussMaxRadius + 1];  // w[d] = weight at distance d (kernel is symmetric)
};

// ---- mbarrier / bulk-copy PTX ------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
// The same wait for callers whose sibling CTAs on the SM have work to issue meanwhile: the thread
// may be suspended up to `hint_ns` per probe instead of re-probing at once.
__device__ __forceinline__ void mbar_wait_suspend(uint64_t *bar, uint32_t parity, uint32_t hint_ns)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP_S:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra WAIT_DONE_S;\n"
        "bra WAIT_LOOP_S;\n"
        "WAIT_DONE_S:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity), "r"(hint_ns)
        : "memory");
}
// The wait for role hand-offs that are expected to block for a while: a non-blocking probe, then a
// plain timed sleep.  try_wait with a suspend hint compiles to TRYWAIT + NANOSLEEP.SYNCS, which any
// mbarrier traffic of the CTA wakes up again (11 wake-ups per wait measured); every probe is issue
// slots and energy, and the sustained Gaussian is power-capped.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity, uint32_t sleep_ns)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP_Z:\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE_Z;\n"
        "nanosleep.u32 %2;\n"
        "bra WAIT_LOOP_Z;\n"
        "WAIT_DONE_Z:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity), "r"(sleep_ns)
        : "memory");
}
// Either of the two, chosen by a launch parameter (bit 30: suspended try_wait with the hint in the low bits)
__device__ __forceinline__ void mbar_wait_cfg(uint64_t *bar, uint32_t parity, uint32_t cfg)
{
    if (cfg & 0x40000000u) mbar_wait_suspend(bar, parity, cfg & 0x3fffffffu);
    else mbar_wait_sleep(bar, parity, cfg);
}
// global -> shared 1-D bulk copy through the TMA unit; bytes % 16 == 0, both sides 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
// the same with both shared-memory operands already 32-bit shared addresses (issue loops step them)
__device__ __forceinline__ void bulk_g2s_raw(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar_smem)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar_smem)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}


}  // namespace mpk
