// Separable Gaussian, fp32 HWC, streaming kernel, warp-specialised (v3).
//
// Same decomposition as gaussian_stream.cuh -- work item = (image, 640-float
// column strip, row chunk), rows filtered horizontally into shared memory, columns
// accumulated in registers -- but the two passes run on different warps and are
// decoupled by mbarrier rings instead of a CTA-wide barrier per step:
//
//   warps 0..9   ROW warps.  Warp q owns row q of every 10-row group.  It runs its
//                own private TMA pipeline: lane 0 issues one 1-D cp.async.bulk
//                (UBLKCP) per row into a 3-slot ring nobody else touches, waits on the
//                slot's mbarrier, filters the row (packed FFMA2, two accumulator sets
//                so that even- and odd-offset taps both read the register pairs
//                exactly as LDS.128 delivered them), stores 640 filtered floats into
//                the group's slot of the hand-off ring and arrives on full[group].
//   warps 10..19 COLUMN warps.  Thread t owns float columns 2t, 2t+1 and the 2R+1
//                live output rows of each, packed in registers.  Per group: wait
//                full[group], 10 x (LDS.64, 2R FFMA2 + 1 FMUL2 that accumulate AND slide
//                the window -- A[j-1] = fma(w, v, A[j]) -- one coalesced 8-byte
//                streaming store), arrive on empty[group].
//
// One CTA of 640 threads per SM (the register file splits 102 per thread, which
// both roles fit), 3 groups in flight between the roles, 3 rows in flight per ROW
// warp from HBM.  Per 10-row group the FMA pipe needs ~2360 cycles per SMSP and the
// issue port ~1600, so the pipe -- not issue, not shared memory, not a barrier -- is
// the limiter; see DESIGN.md for the arithmetic and profiles/ for the measurement.
#pragma once
#include "gaussian_stream.cuh"

namespace mpk {

// ROW warps = rows per group.  10:10 with the COLUMN warps is what the register file allows: 21
// warps put 6 on one SM sub-partition (16 K registers each), which caps a thread at 80 registers
// and spills the row pass (measured: 11:10 at 80 registers is 8 % slower than 10:10 at 91).
constexpr int kWsRows = 10;
constexpr int kWsRowWarps = kWsRows;
constexpr int kWsColWarps = kGsTW / 64;       // 10: 320 threads x 2 columns = 640 floats
constexpr int kWsThreads = 32 * (kWsRowWarps + kWsColWarps);
constexpr int kWsInSlots = 4;                 // rows in flight per ROW warp
constexpr int kWsGroups = 4;                  // filtered groups in flight between the roles
constexpr uint32_t kWsSleepNs = 100;          // sleep between probes of a blocked hand-off wait (MMA flavour)

// The same numbers per kernel flavour: the FMA-column kernel (this file) splits the CTA 10:10, the
// tensor-core-column kernel (gaussian_stream_mma.cuh) 12:8 on warpgroup boundaries (setmaxnreg).
template <bool MMA>
struct WsK {
    static constexpr int rows = MMA ? 12 : kWsRows;          // rows per group = ROW warps
    static constexpr int row_warps = rows;
    static constexpr int col_warps = MMA ? 8 : kWsColWarps;
    static constexpr int in_slots = MMA ? 2 : kWsInSlots;   // MMA: the third slot's room stages the output
    static constexpr int groups = kWsGroups;
    static constexpr int threads = 32 * (row_warps + col_warps);
};

template <int C, int R>
struct WsGeom {
    static constexpr int HALO = GsGeom<C, R>::HALO;
    static constexpr int ROW = GsGeom<C, R>::ROW;
    static constexpr int SLOT = GsGeom<C, R>::SLOT;
    static constexpr size_t IN_BYTES = (size_t)kWsRowWarps * kWsInSlots * SLOT * 4;
    static constexpr size_t H_BYTES = (size_t)kWsGroups * kWsRows * kGsTW * 4;
    static constexpr int N_BARS = kWsRowWarps * kWsInSlots + 2 * kWsGroups;
    static constexpr size_t PROG_BYTES = (size_t)(kWsThreads / 32) * sizeof(PwSmem);   // one fused program per warp
    static constexpr size_t SMEM = IN_BYTES + H_BYTES + 8 * N_BARS + 64 + PROG_BYTES;
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

// ---- row pass, two accumulator sets ---------------------------------------------
// A[m] = outputs (2m, 2m+1) fed by taps with an even offset k*C; B[m] = outputs
// (2m-1, 2m) fed by taps with an odd offset.  Either way the operand is the aligned
// pair (x[2j], x[2j+1]) as loaded.  out[i] = A-part + B-part.
// `W` is a small accessor whose operator()(d) yields the packed weight pair of distance d from the
// constant bank (WsOneSet: the launch's single set; WsSetOf: the work item's image's set).
struct WsOneSet {
    const GaussStreamParams &p;
    __device__ __forceinline__ uint64_t operator()(int d) const { return p.ww[d]; }
};
struct WsSetOf {
    const GaussWeightSets &ws;
    int set;
    __device__ __forceinline__ uint64_t operator()(int d) const { return ws.ww[set][d]; }
};

template <int C, int R, typename W>
__device__ __forceinline__ void ws_row_pass(const float *__restrict__ win, float (&out)[kGsPH], const W &w)
{
    constexpr int HALO = GsGeom<C, R>::HALO;
    constexpr int NV = (kGsPH + 2 * HALO) / 4;
    constexpr int NB = kGsPH / 2 + 1;
    uint64_t A[kGsPH / 2], B[NB];
#pragma unroll
    for (int i = 0; i < kGsPH / 2; ++i) A[i] = 0ull;
#pragma unroll
    for (int i = 0; i < NB; ++i) B[i] = 0ull;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const ulonglong2 ld = *reinterpret_cast<const ulonglong2 *>(win + 4 * v);
        const uint64_t e[2] = {ld.x, ld.y};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int j = 2 * v + u;  // this pair holds window offsets 2j, 2j+1
#pragma unroll
            for (int k = -R; k <= R; ++k) {
                const int kc = k * C;
                const int i0 = 2 * j - HALO - kc;  // output fed by the pair's low half
                if ((kc & 1) == 0) {
                    if (i0 >= 0 && i0 < kGsPH) A[i0 / 2] = ffma2(w(k < 0 ? -k : k), e[u], A[i0 / 2]);
                } else {
                    // i0 is odd: the pair feeds outputs (i0, i0 + 1) = B[(i0 + 1) / 2]
                    if (i0 >= -1 && i0 < kGsPH) B[(i0 + 1) / 2] = ffma2(w(k < 0 ? -k : k), e[u], B[(i0 + 1) / 2]);
                }
            }
        }
    }
    constexpr bool any_odd = (C & 1) != 0;
#pragma unroll
    for (int m = 0; m < kGsPH / 2; ++m) {
        float a_lo, a_hi;
        unpack2(A[m], a_lo, a_hi);
        if (any_odd) {
            float b0_lo, b0_hi, b1_lo, b1_hi;
            unpack2(B[m], b0_lo, b0_hi);
            unpack2(B[m + 1], b1_lo, b1_hi);
            out[2 * m] = a_lo + b0_hi;
            out[2 * m + 1] = a_hi + b1_lo;
        } else {
            out[2 * m] = a_lo;
            out[2 * m + 1] = a_hi;
        }
    }
}

// Steps (10-row groups) a work item takes.  The tensor-core column pass (gaussian_stream_mma.cuh)
// consumes the hand-off ring in 8-row chunks and finishes an output block NCH - 1 chunks after its
// first row arrived, so there an item is padded by 7 rows and rounded up to a whole turn of the
// 40-row ring (4 groups = 5 chunks): every item starts on ring row 0 with all barriers in phase.
template <bool MMA>
__device__ __forceinline__ int ws_steps(int n_rows)
{
    using K = WsK<MMA>;
    if (!MMA) return (n_rows + K::rows - 1) / K::rows;
    return ((n_rows + 7 + K::rows - 1) / K::rows + 3) & ~3;
}

// The ROW-warp role: private TMA ring -> horizontal filter -> hand-off ring (row pitch PITCH floats).
template <int C, int R, bool SETS, int PITCH, bool MMA>
__device__ __forceinline__ void ws_row_role(const GaussStreamParams &p, const GaussWeightSets &ws, float *s_in,
                                            float *s_h, uint64_t *in_full, uint64_t *h_full, uint64_t *h_empty,
                                            int warp, int lane, PwSmem *s_prog)
{
    PwSmem *my_prog = s_prog + warp;   // this warp's copy of the current image's "before the blur" program
    using G = WsGeom<C, R>;
    using K = WsK<MMA>;
    const long n_items = gs_item_count(p);
    // =============================================================== ROW warp
    float *my_in = s_in + (size_t)warp * K::in_slots * G::SLOT;
    uint64_t *my_full = in_full + warp * K::in_slots;
    uint32_t loads = 0;   // rows issued so far by this warp (slot = loads % 3, parity from the count)
    uint32_t takes = 0;   // rows consumed so far
    uint32_t group = 0;   // groups produced so far (ring slot and parity)

    for (long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const GsItem it = gs_item(p, item);
        const int img = it.img, strip = it.strip;
        const float *__restrict__ src = p.in_tab ? p.in_tab[img] : p.in + (size_t)img * p.image_stride;
        const WsSetOf w_img{ws, img % kGsMaxSets};
        const WsOneSet w_one{p};
        const int x0 = strip * kGsTW;
        const int y0 = it.y0, y1 = it.y1;
        const int r_begin = y0 - R;
        const int n_rows = (y1 - y0) + 2 * R;
        const int n_steps = ws_steps<MMA>(n_rows);
        const int gx_start = x0 - G::HALO;
        const int lo = gx_start < 0 ? -gx_start : 0;
        const int hi = min(G::ROW, p.row_elems - gx_start);
        const uint32_t row_bytes = (uint32_t)(hi - lo) * 4u;
        // this image's "before the blur" program (null or empty: none) and the channel of ring column 0
        int n_pre = 0;
        if (SETS && p.pw_tab) {   // staged in this warp's shared-memory slot: it is applied to every row of the item
            pw_smem_load(my_prog, p.pw_tab + (size_t)img * p.pw_stride, lane);
            n_pre = my_prog->n;
        }
        const int ch_start = ((gx_start % C) + C) % C;
        const int pre_ch0 = (ch_start + lo + 4 * lane) % C;   // channel of the first sample of this lane's first vector

        // every load this warp issued has been consumed: its ring is quiescent
        if (lo > 0 || hi < G::ROW) {
            for (int i = lane; i < K::in_slots * G::SLOT; i += 32) {
                const int col = i % G::SLOT;
                if (col < lo || col >= hi) my_in[i] = 0.f;
            }
            fence_proxy_async();
        }
        __syncwarp();

        // This warp's rows are r_begin + warp + 10*step.  Steps [live_lo, live_hi) are the ones
        // whose row lies inside the image and the item; everything per step is then a running
        // pointer and two compares.
        const int r_first = r_begin + warp;
        const int r_end = min(p.height, r_begin + n_rows);
        int live_lo = r_first < 0 ? (-r_first + K::rows - 1) / K::rows : 0;
        int live_hi = r_end > r_first ? (r_end - r_first + K::rows - 1) / K::rows : 0;
        if (live_hi > n_steps) live_hi = n_steps;
        if (live_lo > live_hi) live_lo = live_hi;
        const float *gptr = src + (long)(r_first + live_lo * K::rows) * p.row_elems + gx_start + lo;
        const long gstep = (long)K::rows * p.row_elems;
        int next_issue = live_lo;  // next live step to issue

        // Rows that are not whole vectors (W * C % 4 != 0; *_sets kernels only): a row starts `sh` floats
        // past a 16-byte boundary, sh = (row * (W * C % 4)) % 4.  The copy then starts at that boundary
        // and covers the enclosing aligned span (it stays inside the allocation: the pool rounds sizes
        // up to 16 bytes); the row is shifted into place after it has landed.
        const int mis = SETS ? (p.row_elems & 3) : 0;
        int row_issue = r_first + live_lo * K::rows;   // image row of the next copy / of the next row consumed
        int row_take = row_issue;
        auto issue_next = [&]() {  // all lanes keep the counters; lane 0 talks to the TMA unit
            const uint32_t slot = loads % K::in_slots;
            if (lane == 0) {
                if (SETS && mis) {
                    const int sh = (row_issue * mis) & 3;
                    const uint32_t bytes = ((uint32_t)(hi - lo + sh) * 4u + 15u) & ~15u;
                    mbar_expect_tx(&my_full[slot], bytes);
                    bulk_g2s(my_in + (size_t)slot * G::SLOT + lo, gptr - sh, bytes, &my_full[slot]);
                } else {
                    mbar_expect_tx(&my_full[slot], row_bytes);
                    bulk_g2s(my_in + (size_t)slot * G::SLOT + lo, gptr, row_bytes, &my_full[slot]);
                }
            }
            gptr += gstep;
            row_issue += K::rows;
            ++loads;
            ++next_issue;
        };
        // prologue: up to K::in_slots - 1 rows in flight before the first one is consumed
        for (int s = 0; s < K::in_slots - 1 && next_issue < live_hi; ++s) issue_next();

        float *hbase = s_h + (size_t)warp * PITCH + lane * kGsPH;
        for (int step = 0; step < n_steps; ++step) {
            const bool live = step >= live_lo && step < live_hi;
            float out[kGsPH];
            if (live) {
                // the slot consumed in the previous live step is free again: keep the ring full
                if (next_issue < live_hi) issue_next();
                const uint32_t slot = takes % K::in_slots;
                if (MMA) mbar_wait_cfg(&my_full[slot], (takes / K::in_slots) & 1u, p.wait_ns[0]);
                else mbar_wait(&my_full[slot], (takes / K::in_slots) & 1u);
                ++takes;
                if (SETS && mis) {
                    // shift the landed row sh floats down into place: block b of the destination is the
                    // tail of landed block b and the head of block b + 1.  All lanes load, then all store
                    // (lane l + 1 rewrites what lane l reads); what lies beyond `hi` becomes zero padding.
                    const int sh = (row_take * mis) & 3;
                    float *row = my_in + (size_t)slot * G::SLOT;
                    constexpr int NQ = (G::SLOT / 4 + 31) / 32;
#pragma unroll 1
                    for (int j = 0; j < NQ; ++j) {
                        const int i = lo + 4 * (lane + 32 * j);
                        float o[4] = {0.f, 0.f, 0.f, 0.f};
                        if (i < hi) {
                            const float4 v0 = *reinterpret_cast<const float4 *>(row + i);
                            const float4 v1 = *reinterpret_cast<const float4 *>(row + i + 4);   // within SLOT: i + 7 < hi + 7 <= ROW + 4 + 3
                            const float w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float pick = sh == 0 ? w[e] : (sh == 1 ? w[e + 1] : (sh == 2 ? w[e + 2] : w[e + 3]));
                                o[e] = i + e < hi ? pick : 0.f;
                            }
                        }
                        __syncwarp();
                        if (i < G::SLOT) *reinterpret_cast<float4 *>(row + i) = make_float4(o[0], o[1], o[2], o[3]);
                        __syncwarp();
                    }
                    fence_proxy_async();   // the slot's next writer is the TMA unit
                    __syncwarp();
                }
                row_take += K::rows;
                if (SETS && n_pre) {
                    // fused pointwise ops in front of the blur: rewrite the landed row in place, each
                    // sample once (the lanes' filter windows overlap 4.6x, so doing it on the window
                    // registers would repeat the work).  Only the in-image part [lo, hi): what lies
                    // outside is the blur's zero padding of the *transformed* image and stays zero.
                    float *row = my_in + (size_t)slot * G::SLOT;
                    constexpr int NQ = (G::ROW / 4 + 31) / 32;   // vectors per lane (6 for a 712-float row)
#pragma unroll
                    for (int j0 = 0; j0 < NQ; j0 += 3) {         // three vectors at a time: 12 registers
                        float v[3][4];
#pragma unroll
                        for (int u = 0; u < 3; ++u) {
                            const int i = lo + 4 * (lane + 32 * (j0 + u));
                            float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (j0 + u < NQ && i < hi) t4 = *reinterpret_cast<const float4 *>(row + i);
                            v[u][0] = t4.x; v[u][1] = t4.y; v[u][2] = t4.z; v[u][3] = t4.w;
                        }
                        for (int k = 0; k < n_pre; ++k) {
                            const PwOp op = pw_smem_op(*my_prog, k);
#pragma unroll
                            for (int u = 0; u < 3; ++u)   // vector j starts 128 j floats after vector 0: 128 = 2 (mod 3), 0 (mod 4)
                                pw_apply_op_tile<C, 4>(op, v[u], pw_channel<C>(pre_ch0, C == 3 ? (2 * (j0 + u)) % 3 : 0));
                        }
#pragma unroll
                        for (int u = 0; u < 3; ++u) {
                            const int i = lo + 4 * (lane + 32 * (j0 + u));
                            if (j0 + u < NQ && i < hi) {
                                if (SETS && mis) {   // a row may end inside a vector: what lies beyond stays zero padding
#pragma unroll
                                    for (int e = 1; e < 4; ++e)
                                        if (i + e >= hi) v[u][e] = 0.f;
                                }
                                *reinterpret_cast<float4 *>(row + i) = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
                            }
                        }
                    }
                    fence_proxy_async();   // the slot's next writer is the TMA unit
                    __syncwarp();
                }
                if (SETS) ws_row_pass<C, R>(my_in + (size_t)slot * G::SLOT + lane * kGsPH, out, w_img);
                else ws_row_pass<C, R>(my_in + (size_t)slot * G::SLOT + lane * kGsPH, out, w_one);
            } else {
#pragma unroll
                for (int i = 0; i < kGsPH; ++i) out[i] = 0.f;
            }
            // hand-off ring: wait until the COLUMN warps have drained this group slot
            const uint32_t gs = group % K::groups;
            if (MMA) mbar_wait_cfg(&h_empty[gs], ((group / K::groups) & 1u) ^ 1u, p.wait_ns[1]);
            else mbar_wait(&h_empty[gs], ((group / K::groups) & 1u) ^ 1u);
            float *hrow = hbase + (size_t)gs * (K::rows * PITCH);
#pragma unroll
            for (int v = 0; v < kGsPH / 4; ++v)
                *reinterpret_cast<float4 *>(hrow + 4 * v) =
                    make_float4(out[4 * v], out[4 * v + 1], out[4 * v + 2], out[4 * v + 3]);
            __syncwarp();  // all lanes' stores and window reads are done
            if (lane == 0) mbar_arrive(&h_full[gs]);
            ++group;
        }
    }
}

template <int C, int R, bool SETS>
__device__ __forceinline__ void ws_body(const GaussStreamParams &p, const GaussWeightSets &ws)
{
    using G = WsGeom<C, R>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_in = reinterpret_cast<float *>(smem_raw);                // [10 warps][3 slots][ROW]
    float *s_h = reinterpret_cast<float *>(smem_raw + G::IN_BYTES);    // [3 groups][10 rows][640]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + G::IN_BYTES + G::H_BYTES);
    uint64_t *in_full = bars;                                        // [10][3]
    uint64_t *h_full = bars + kWsRowWarps * kWsInSlots;               // [3]
    uint64_t *h_empty = h_full + kWsGroups;                           // [3]
    PwSmem *s_prog = reinterpret_cast<PwSmem *>(smem_raw + G::IN_BYTES + G::H_BYTES + 8 * G::N_BARS + 64);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < kWsRowWarps * kWsInSlots; ++i) mbar_init(&in_full[i], 1);
        for (int g = 0; g < kWsGroups; ++g) {
            mbar_init(&h_full[g], kWsRowWarps);
            mbar_init(&h_empty[g], kWsColWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();  // the only CTA-wide barrier

    const long n_items = gs_item_count(p);

    if (warp < kWsRowWarps) {
        ws_row_role<C, R, SETS, kGsTW, false>(p, ws, s_in, s_h, in_full, h_full, h_empty, warp, lane, s_prog);
    } else {
        // ============================================================ COLUMN warp
        // A[j] is the partial sum of output row (r - R + j) when filtered row r arrives.  Row r
        // adds w[|R - j|] * h[r] to it and the window slides by one: both happen in ONE
        // instruction per accumulator, A[j-1] = fma(w, v, A[j]) -- the destination is the
        // neighbour, so nothing rotates, no phase exists, and every row runs the same ~30
        // instructions.  A[0] + w[R] * v is the finished output row r - R.
        const int vt = tid - 32 * kWsRowWarps;  // 0..319
        uint64_t A[2 * R];
#pragma unroll
        for (int i = 0; i < 2 * R; ++i) A[i] = 0ull;
        uint32_t group = 0;

        for (long item = blockIdx.x; item < n_items; item += gridDim.x) {
            const GsItem it = gs_item(p, item);
            const int img = it.img, strip = it.strip;
            const int gx = strip * kGsTW + 2 * vt;
            const int y0 = it.y0, y1 = it.y1;
            const int n_rows = (y1 - y0) + 2 * R;
            const int n_steps = (n_rows + kWsRows - 1) / kWsRows;
            // rows [y0, y1) of columns gx, gx+1; the first filtered row completes output row y0 - 2R
            const unsigned n_valid = gx < p.row_elems ? (unsigned)(y1 - y0) : 0u;
            float *base = p.out_tab ? p.out_tab[img] : p.out + (size_t)img * p.image_stride;
            const int set = SETS ? img % kGsMaxSets : 0;
            auto w = [&](int d) -> uint64_t { return SETS ? ws.ww[set][d] : p.ww[d]; };
            float pb = 0.f, plo = 0.f, phi = 0.f;   // "after the blur": out = min(max(v + pb, plo), phi)
            bool has_post = false;
            if (SETS && p.pw_tab) {
                const PwProgram *pp = p.pw_tab + (size_t)img * p.pw_stride + 1;
                if (__ldg(&pp->n) != 0) {
                    has_post = true;
                    pb = __ldg(&pp->ops[0].a);
                    plo = __ldg(&pp->ops[0].b);
                    phi = __ldg(&pp->ops[0].c);
                }
            }
            float *optr = base + ((long)y0 - 2 * R) * p.row_elems + gx;  // only dereferenced when valid
            unsigned rel = (unsigned)(-2 * R);                            // output row - y0, wraps below 0

            for (int step = 0; step < n_steps; ++step) {
                const uint32_t gs = group % kWsGroups;
                mbar_wait(&h_full[gs], (group / kWsGroups) & 1u);
                const float *hrow = s_h + (size_t)gs * kWsRows * kGsTW + 2 * vt;
#pragma unroll
                for (int q = 0; q < kWsRows; ++q) {
                    const uint64_t v = *reinterpret_cast<const uint64_t *>(hrow + q * kGsTW);
                    const uint64_t o = ffma2(w(R), v, A[0]);
#pragma unroll
                    for (int j = 1; j < 2 * R; ++j) A[j - 1] = ffma2(w(j < R ? R - j : j - R), v, A[j]);
                    A[2 * R - 1] = fmul2(w(R), v);
                    if (rel < n_valid) {
                        float o2[2];
                        unpack2(o, o2[0], o2[1]);
                        if (SETS && has_post) {
                            o2[0] = fminf(fmaxf(o2[0] + pb, plo), phi);
                            o2[1] = fminf(fmaxf(o2[1] + pb, plo), phi);
                        }
                        if (SETS && (p.row_elems & 1)) {   // odd rows: the pair may be misaligned or straddle the row end
                            __stcs(optr, o2[0]);
                            if (gx + 1 < p.row_elems) __stcs(optr + 1, o2[1]);
                        } else {
                            __stcs(reinterpret_cast<float2 *>(optr), make_float2(o2[0], o2[1]));
                        }
                    }
                    ++rel;
                    optr += p.row_elems;
                }
                __syncwarp();  // every lane has read the group's rows
                if (lane == 0) mbar_arrive(&h_empty[gs]);
                ++group;
            }
        }
    }
}

template <int C, int R>
__global__ void __launch_bounds__(kWsThreads, 1)
gauss_stream_ws_kernel(const __grid_constant__ GaussStreamParams p)
{
    // one weight set for the whole launch; the reference below is never read
    ws_body<C, R, false>(p, *reinterpret_cast<const GaussWeightSets *>(&p));
}

// Same kernel, weights per image (see GaussWeightSets).
template <int C, int R>
__global__ void __launch_bounds__(kWsThreads, 1)
gauss_stream_ws_sets_kernel(const __grid_constant__ GaussStreamParams p, const __grid_constant__ GaussWeightSets ws)
{
    ws_body<C, R, true>(p, ws);
}

}  // namespace mpk
