// Separable Gaussian, fp32 HWC, streaming kernel (v4): the column pass on the tensor cores.
// Reference: the two-pass choreography of src/millipyde_image.cpp:725-789 (_gaussian_greyscale) and
// its kernels :146-244, under the oracle rule (scipy weights, radius int(8 sigma + 0.5)).
//
// gaussian_stream_ws.cuh is bound by the fp32 FMA pipe: 2 x (2R+1) FMAs per sample put the pipe and
// HBM at the same throughput, and the pipe never runs at 100 %.  Here the ROW warps (unchanged:
// private TMA ring, FFMA2 horizontal filter, hand-off ring) keep the FMA pipe to themselves and the
// COLUMN warps run the vertical filter as small banded matrix products on the otherwise idle tensor
// pipe (mma.sync.m16n8k8 tf32, SASS HMMA.1688.F32.TF32), which halves the FMA-pipe load.
//
// The product.  Filtered rows arrive in 8-row chunks (chunk s = rows f in [8s, 8s+8) of the item);
// output rows are produced in 8-row blocks (block b = rows o in [8b, 8b+8)); out[o] = sum_f
// w[|f - o - R|] F[f] touches the NCH = (7 + 2R)/8 + 1 chunks b .. b+NCH-1.  One MMA computes, for a
// 16-column tile,
//     D[16 columns x 8 output rows] += A[16 columns x 8 chunk rows] . B_j[8 chunk rows x 8 output rows]
// with B_j[k][n] = w[|8j + k - n - R|] (zero outside the support), j = s - b = the block's age.
// A chunk is loaded ONCE and multiplied into the NCH live blocks; like the FMA version's sliding
// accumulators the destination of the first MMA is the neighbour (acc[j+1] = A . B_j + acc[j]), so
// nothing rotates.  The block of age NCH-1 is complete: it is staged in shared memory and leaves as
// TMA bulk stores (cp.async.bulk shared -> global), because accumulator fragments are row-scattered
// and a direct STG.64 costs 4x the L1 data-pipe wavefronts.
//
// Precision.  tf32 carries 11 significant bits, so both operands are split, x = hi + lo with
// hi = x & 0xffffe000 (exact), and the product is hi.hi (one tf32 MMA) + lo.hi + hi.lo.  The two
// correction products are 2^-10 of the result and need 11 bits themselves, which fp16 has: they are
// ONE m16n8k16 fp16 MMA whose K dimension is the concatenation [lo(A) | hi(A)] x [hi(B) ; lo(B)].
// Everything accumulates in fp32.  Dropped: lo.lo and the rounding of the lo parts, each < 2^-21
// relative -- ~3e-7 absolute on [0,1] images against the fp64 oracle (contract 1e-5).  The fp16
// operands limit the kernel to |sample| < 65504 (an fp32 *image*; the FMA-pipe kernel has no limit).
//
// Fragment mapping (g = lane >> 2, t = lane & 3).  MMA row m <-> tile column: m = g -> column 2g,
// m = g + 8 -> column 2g + 1, so a thread's A operands (a0, a1) and (a2, a3) are two LDS.64 from ring
// rows t and t + 4, and its results (c0, c2) / (c1, c3) are column pairs of output rows 2t / 2t + 1.
// The ring's row pitch is 648 floats (= 8 mod 32): the LDS.64 of a half-warp hit 32 distinct banks.
#pragma once
#include "gaussian_stream_ws.cuh"

namespace mpk {

using MmK = WsK<true>;                          // 12 ROW warps (12-row groups), 8 COLUMN warps
constexpr int kMmPitch = kGsTW + 8;            // hand-off ring row pitch in floats
constexpr int kMmRing = MmK::groups * MmK::rows;   // 48 rows = 4 groups of 12 = 6 chunks of 8
constexpr int kMmTiles = kGsTW / 16 / MmK::col_warps;  // 16-column tiles per COLUMN warp (5)
constexpr int kMmThreads = MmK::threads;       // 640
// Output staging: 8 rows x 80 floats per COLUMN warp at a pitch of 88 floats (= 24 mod 32: rows
// t and 4+t of the four t's of a half-warp start 8 banks apart, so its 8-byte stores do not conflict).
constexpr int kMmStagePitch = 88;
// Registers: the kernel is compiled for 96 per thread (640 threads) and that allocation is the
// CTA's pool.  The 12 ROW warps (warpgroups 0-2) hand 16 each back (setmaxnreg.dec 80), the 8 COLUMN
// warps (warpgroups 3-4) take them (setmaxnreg.inc 120): 12*80 + 8*120 = 20*96.
constexpr int kMmRowRegs = 80;
constexpr int kMmColRegs = 120;

template <int C, int R>
struct MmGeom {
    static constexpr int NCH = (7 + 2 * R) / 8 + 1;   // chunks an output block spans
    static constexpr size_t IN_BYTES = (size_t)MmK::row_warps * MmK::in_slots * GsGeom<C, R>::SLOT * 4;
    static constexpr size_t H_BYTES = (size_t)kMmRing * kMmPitch * 4;
    static constexpr size_t STAGE_BYTES = (size_t)MmK::col_warps * 8 * kMmStagePitch * 4;
    static constexpr int N_BARS = MmK::row_warps * MmK::in_slots + 2 * MmK::groups;
    static constexpr size_t PROG_BYTES = (size_t)(MmK::threads / 32) * sizeof(PwSmem);   // one fused program per warp
    static constexpr size_t SMEM = IN_BYTES + H_BYTES + STAGE_BYTES + 8 * N_BARS + 64 + PROG_BYTES;
};

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1,
                                         const float (&c)[4])
{
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%10,%11,%12,%13};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c[0]), "f"(c[1]), "f"(c[2]),
          "f"(c[3]));
}

// m16n8k16, fp16 operands, fp32 accumulate (SASS HMMA.16816.F32): the two correction products of a
// (tile, block) pair in one instruction, K slots 0..7 = lo(A) x hi(B), 8..15 = hi(A) x lo(B).
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1,
                                        const float (&c)[4])
{
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%10,%11,%12,%13};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c[0]), "f"(c[1]), "f"(c[2]),
          "f"(c[3]));
}
// (x, y) -> packed f16x2 with x in the low half: the K slot pair (2t, 2t + 1) of one MMA operand
__device__ __forceinline__ uint32_t pack_f16(float x, float y)
{
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(y), "f"(x));
    return r;
}
// shared -> global 1-D bulk copy through the TMA unit (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr(src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void sts_f2(float *p, float x, float y)
{
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(smem_addr(p)), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ float2 lds_f2(const float *p)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(smem_addr(p)));
    return v;
}

template <int C, int R, bool SETS>
__device__ __forceinline__ void mm_body(const GaussStreamParams &p, const GaussWeightSets &ws)
{
    using G = MmGeom<C, R>;
    constexpr int NCH = G::NCH;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_in = reinterpret_cast<float *>(smem_raw);                 // [12 warps][3 slots][ROW]
    float *s_h = reinterpret_cast<float *>(smem_raw + G::IN_BYTES);    // [48 rows][648]
    float *s_stage = reinterpret_cast<float *>(smem_raw + G::IN_BYTES + G::H_BYTES);   // [8 warps][8 rows][88]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + G::IN_BYTES + G::H_BYTES + G::STAGE_BYTES);
    uint64_t *in_full = bars;
    uint64_t *h_full = bars + MmK::row_warps * MmK::in_slots;
    uint64_t *h_empty = h_full + MmK::groups;
    PwSmem *s_prog = reinterpret_cast<PwSmem *>(smem_raw + G::IN_BYTES + G::H_BYTES + G::STAGE_BYTES + 8 * G::N_BARS + 64);

    // the warp index is read from lane 0 so that the compiler knows it is warp-uniform: ring slots,
    // barrier and bulk-copy addresses then live in uniform registers (UBLKCP without R2UR loops)
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < MmK::row_warps * MmK::in_slots; ++i) mbar_init(&in_full[i], 1);
        for (int g = 0; g < MmK::groups; ++g) {
            mbar_init(&h_full[g], MmK::row_warps);
            mbar_init(&h_empty[g], MmK::col_warps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();  // the only CTA-wide barrier

    if (warp < MmK::row_warps) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kMmRowRegs));
        ws_row_role<C, R, SETS, kMmPitch, true>(p, ws, s_in, s_h, in_full, h_full, h_empty, warp, lane, s_prog);
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kMmColRegs));

    // ================================================================ COLUMN warp
    const long n_items = gs_item_count(p);
    const int wc = warp - MmK::row_warps;   // 0..7: columns [80 wc, 80 wc + 80) of the strip
    const int g = lane >> 2, t = lane & 3;
    const float *ring = s_h + t * kMmPitch + wc * (16 * kMmTiles) + 2 * g;
    float *stage = s_stage + wc * (8 * kMmStagePitch);              // this warp's staging rows
    float *my_stage = stage + t * kMmStagePitch + 2 * g;           // block row 2t (row 2t+1: 4 rows further)

    uint32_t waited = 0;     // groups of the hand-off ring waited for so far (all items)
    uint32_t released = 0;   // groups handed back so far

    for (long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const GsItem it = gs_item(p, item);
        const int img = it.img, strip = it.strip;
        const int y0 = it.y0, y1 = it.y1;
        const int n_rows = (y1 - y0) + 2 * R;
        const int n_chunks8 = ws_steps<true>(n_rows) / 4 * (kMmRing / 8);
        const unsigned n_valid = (unsigned)(y1 - y0);
        float *base = p.out_tab ? p.out_tab[img] : p.out + (size_t)img * p.image_stride;

        // B fragments of this item's weight set.  tf32 product: bh[j] = hi(B_j[t][g]), hi(B_j[t+4][g]);
        // fp16 correction product: bc[j][0] = (B_j[t][g], B_j[t+4][g]) against lo(A), bc[j][1] = the
        // lo parts of the same two weights against hi(A).
        uint32_t bh[NCH][2], bc[NCH][2];
        {
            const int set = SETS ? img % kGsMaxSets : 0;
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                float w[2], wl[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    int d = 8 * j + t + 4 * h - g - R;
                    d = d < 0 ? -d : d;
                    w[h] = 0.f;
                    if (d <= R) w[h] = SETS ? __uint_as_float((uint32_t)ws.ww[set][d]) : p.w[d];
                    bh[j][h] = __float_as_uint(w[h]) & 0xffffe000u;
                    wl[h] = w[h] - __uint_as_float(bh[j][h]);
                    asm volatile("" : "+r"(bh[j][h]));   // keep it: re-deriving it per chunk costs an indexed LDC
                }
                bc[j][0] = pack_f16(w[0], w[1]);
                bc[j][1] = pack_f16(wl[0], wl[1]);
                asm volatile("" : "+r"(bc[j][0]), "+r"(bc[j][1]));
            }
        }

        // acc[tile][j - 1] = partial sums of the block that has seen j chunks, j = 1 .. NCH-1
        float acc[kMmTiles][NCH - 1][4];
#pragma unroll
        for (int q = 0; q < kMmTiles; ++q)
#pragma unroll
            for (int j = 0; j < NCH - 1; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[q][j][e] = 0.f;

        // "After the blur": one add-and-clamp, out = min(max(v + pb, plo), phi) -- what brightness, clip and
        // the clamps of colorize / multiply compose to (their per-channel factors run BEFORE the blur: it
        // is linear).  Three straight-line instructions per accumulator, so they interleave with the
        // MMAs; a general program here (a loop over ops between the MMAs and the staging stores) put
        // the role's chunk latency over its budget and halved the kernel.
        float pb = 0.f, plo = 0.f, phi = 0.f;
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
        const uint32_t item_g0 = waited;   // == released: items are whole turns of the ring
        int ring_row = 0;                  // (c % 6) * 8
        // Output.  The block that completes with chunk c is rows ob .. ob+7, ob = 8 (c - NCH + 1).  Its
        // 8 x 80 samples are staged in the warp's own shared-memory rows and leave as eight 320-byte
        // TMA bulk stores (SASS UBLKCP.G.S) issued by lane 0: a direct store of the accumulator
        // fragments touches 4 rows x 32 bytes per half-warp and costs 4x the shared/L1 data-pipe
        // wavefronts.  Thread (g, t) holds block rows 2t and 2t+1; they are staged in rows t and 4+t.
        int ob = -8 * (NCH - 1);
        const int gx0 = strip * kGsTW + wc * (16 * kMmTiles);     // first column of the warp's slice
        const int cols = min(16 * kMmTiles, p.row_elems - gx0);    // <= 0: the slice is outside the image
        float *gdst = base + ((long)y0 + ob) * p.row_elems + gx0;  // block row 0; only used when valid

        for (int c = 0; c < n_chunks8; ++c) {
            const uint32_t need = item_g0 + (uint32_t)(8 * c + 7) / MmK::rows + 1u;
            if (waited < need) {   // the chunk's last row completes at most one more 12-row group
                if (p.col_wait) mbar_wait_suspend(&h_full[waited % MmK::groups], (waited / MmK::groups) & 1u, 1000);
                else mbar_wait_cfg(&h_full[waited % MmK::groups], (waited / MmK::groups) & 1u, p.wait_ns[2]);
                ++waited;
            }
            const float *rp = ring + ring_row * kMmPitch;
            // the chunk's samples of all tiles first: the loads queue behind the ROW warps' window
            // reads on the shared-memory pipe, so they are issued before any of them is needed
            float2 raw[kMmTiles][2];
#pragma unroll
            for (int q = 0; q < kMmTiles; ++q) {
                raw[q][0] = lds_f2(rp + 16 * q);
                raw[q][1] = lds_f2(rp + 4 * kMmPitch + 16 * q);
            }
            // Tiles go through in pairs: the correction MMAs of both tiles, then the tf32 MMAs of both,
            // so the two dependent MMAs of a (tile, block) are 8 tensor instructions apart.
#pragma unroll
            for (int q0 = 0; q0 < kMmTiles; q0 += 2) {
                constexpr int kPair = 2;
                uint32_t ah[kPair][4], ac[kPair][4];
                float nxt[kPair][NCH][4];   // nxt[.][j] = block of age j + 1 after this chunk (NCH-1: complete)
                const float zero[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int u = 0; u < kPair; ++u) {
                    const int q = q0 + u;
                    if (q >= kMmTiles) continue;
                    const float2 v0 = raw[q][0], v1 = raw[q][1];
                    const float a[4] = {v0.x, v0.y, v1.x, v1.y};   // (row t | t+4) x (column 2g | 2g+1)
                    float al[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        ah[u][e] = __float_as_uint(a[e]) & 0xffffe000u;
                        al[e] = a[e] - __uint_as_float(ah[u][e]);
                    }
                    // correction operand: K slots (2t, 2t+1) = lo parts of rows (t, t+4), (2t+8, 2t+9) = hi parts
                    ac[u][0] = pack_f16(al[0], al[2]);
                    ac[u][1] = pack_f16(al[1], al[3]);
                    ac[u][2] = pack_f16(__uint_as_float(ah[u][0]), __uint_as_float(ah[u][2]));
                    ac[u][3] = pack_f16(__uint_as_float(ah[u][1]), __uint_as_float(ah[u][3]));
                }
                // first MMA of every live block: destination is the next age
#pragma unroll
                for (int u = 0; u < kPair; ++u) {
                    const int q = q0 + u;
                    if (q >= kMmTiles) continue;
#pragma unroll
                    for (int j = NCH - 1; j >= 0; --j) {
                        if (j == 0) mma_f16(nxt[u][0], ac[u], bc[0][0], bc[0][1], zero);
                        else mma_f16(nxt[u][j], ac[u], bc[j][0], bc[j][1], acc[q][j - 1]);
                    }
                }
#pragma unroll
                for (int u = 0; u < kPair; ++u) {
                    const int q = q0 + u;
                    if (q >= kMmTiles) continue;
#pragma unroll
                    for (int j = NCH - 1; j >= 0; --j) mma_tf32(nxt[u][j], ah[u], bh[j][0], bh[j][1], nxt[u][j]);
#pragma unroll
                    for (int j = 0; j < NCH - 1; ++j)
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[q][j][e] = nxt[u][j][e];
                }
                if (q0 == 0) {
                    // the previous block's bulk stores have read the staging rows (a chunk ago: no wait
                    // in practice); only lane 0 has groups, the others fall through
                    bulk_wait_read<0>();
                    __syncwarp();
                }
#pragma unroll
                for (int u = 0; u < kPair; ++u) {
                    const int q = q0 + u;
                    if (q >= kMmTiles) continue;
                    if (SETS && has_post) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) nxt[u][NCH - 1][e] = fminf(fmaxf(nxt[u][NCH - 1][e] + pb, plo), phi);
                    }
                    sts_f2(my_stage + 16 * q, nxt[u][NCH - 1][0], nxt[u][NCH - 1][2]);                        // block row 2t
                    sts_f2(my_stage + 4 * kMmStagePitch + 16 * q, nxt[u][NCH - 1][1], nxt[u][NCH - 1][3]);   // block row 2t + 1
                }
            }
            // hand back every group whose 12 rows are now consumed (the MMAs above used every sample)
            ring_row = ring_row == kMmRing - 8 ? 0 : ring_row + 8;
            const uint32_t done = item_g0 + (uint32_t)(8 * (c + 1)) / MmK::rows;
            fence_proxy_async();   // the staged rows are read by the async proxy
            __syncwarp();          // every lane has read the chunk and staged its part of the block
            if (SETS && (p.row_elems & 3)) {
                // rows that are not whole vectors: their starts are not 16-byte aligned, so no bulk
                // stores -- the staged block leaves as scalar stores, a row per 80 lanes' worth
                if (lane == 0)
                    if (released < done) mbar_arrive(&h_empty[released % MmK::groups]);   // 8 rows of 12-row groups: at most one per chunk
                if (cols > 0) {
#pragma unroll 1
                    for (int r = 0; r < 8; ++r) {
                        if ((unsigned)(ob + r) >= n_valid) continue;
                        const float *srow = stage + ((r >> 1) + 4 * (r & 1)) * kMmStagePitch;
                        float *drow = gdst + (long)r * p.row_elems;
                        for (int cc = lane; cc < cols; cc += 32) __stcs(drow + cc, srow[cc]);
                    }
                }
                __syncwarp();   // the staging rows are free again
            } else if (lane == 0) {
                if (released < done) mbar_arrive(&h_empty[released % MmK::groups]);   // 8 rows of 12-row groups: at most one per chunk
                if (cols > 0) {
                    const uint32_t bytes = (uint32_t)cols * 4u;
                    if (ob >= 0 && ob + 8 <= (int)n_valid) {   // the whole block is inside the item
                        float *d = gdst;
#pragma unroll
                        for (int r = 0; r < 8; ++r, d += p.row_elems)
                            bulk_s2g(d, stage + ((r >> 1) + 4 * (r & 1)) * kMmStagePitch, bytes);
                    } else {
#pragma unroll 1
                        for (int r = 0; r < 8; ++r)
                            if ((unsigned)(ob + r) < n_valid)
                                bulk_s2g(gdst + (long)r * p.row_elems,
                                         stage + ((r >> 1) + 4 * (r & 1)) * kMmStagePitch, bytes);
                    }
                }
                bulk_commit();
            }
            released = done;
            ob += 8;
            gdst += 8l * p.row_elems;
        }
    }
    if (lane == 0) bulk_wait_read<0>();   // shared memory stays valid until the last stores have read it
}

template <int C, int R>
__global__ void __launch_bounds__(kMmThreads, 1)
gauss_stream_mma_kernel(const __grid_constant__ GaussStreamParams p)
{
    mm_body<C, R, false>(p, *reinterpret_cast<const GaussWeightSets *>(&p));
}

template <int C, int R>
__global__ void __launch_bounds__(kMmThreads, 1)
gauss_stream_mma_sets_kernel(const __grid_constant__ GaussStreamParams p, const __grid_constant__ GaussWeightSets ws)
{
    mm_body<C, R, true>(p, ws);
}

}  // namespace mpk
