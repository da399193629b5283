"""CPU models of device-side designs that can be checked without a GPU.

Tensor-core Gaussian column pass (millipyde_b200/csrc/kernels/gaussian_stream_mma.cuh):

  * the arithmetic: banded 8-row-chunk x 8-row-block products with the MMA fragment layouts of
    mma.sync.m16n8k8 / m16n8k16, the hi/lo operand split (tf32 product + one fp16 correction
    product whose K dimension concatenates [lo(A) | hi(A)] x [hi(B) ; lo(B)]) and fp32
    accumulation -- against the plain fp64 column filter;
  * the hand-off protocol: ROW warps produce 12-row groups into a 48-row ring, COLUMN warps
    consume 8-row chunks; the `need` / `done` formulas must never let a chunk be read before
    its rows exist nor a ring row be overwritten before its chunk was consumed, and every item
    must end with the ring empty and the barriers in phase.

Skewed transpose (kernels/geometry.cuh, transpose_tma64_kernel): bank-conflict freedom of the
gather and 16-byte alignment of the bulk-copy destinations.

The GPU parity tests (tests/test_ops_gpu.py) check the kernels themselves; these pin the designs
they implement."""
import numpy as np
import pytest

ROWS_PER_GROUP, GROUPS, CHUNK = 12, 4, 8          # WsK<true>::rows, ::groups; chunk rows
RING = ROWS_PER_GROUP * GROUPS                    # 48


def nch(radius):
    return (7 + 2 * radius) // 8 + 1              # MmGeom::NCH


def steps(n_rows):
    """ws_steps<true>: 12-row groups an item takes (padded by 7 rows, whole turns of the ring)."""
    return ((n_rows + 7 + ROWS_PER_GROUP - 1) // ROWS_PER_GROUP + 3) & ~3


# ------------------------------------------------------------------ arithmetic
def tf32_hi(x):
    """hi = x & 0xffffe000 on the fp32 bit pattern (what the kernel feeds the tf32 MMA)."""
    return (np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def mma(a_frag, b_frag, c_frag, k):
    """D = A.B + C for one warp with the PTX fragment layouts (m16n8k8 tf32: k = 8, two A columns
    and one B row per register; m16n8k16 f16: k = 16, register pairs hold K slots (2t, 2t+1)).
    Products and sums in fp32, as the tensor core accumulates."""
    A = np.zeros((16, k), np.float32)
    B = np.zeros((k, 8), np.float32)
    C = np.zeros((16, 8), np.float32)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        if k == 8:
            A[g, t], A[g + 8, t], A[g, t + 4], A[g + 8, t + 4] = a_frag[lane]
            B[t, g], B[t + 4, g] = b_frag[lane]
        else:
            (A[g, 2 * t], A[g, 2 * t + 1]), (A[g + 8, 2 * t], A[g + 8, 2 * t + 1]) = a_frag[lane][0], a_frag[lane][1]
            (A[g, 2 * t + 8], A[g, 2 * t + 9]), (A[g + 8, 2 * t + 8], A[g + 8, 2 * t + 9]) = a_frag[lane][2], a_frag[lane][3]
            (B[2 * t, g], B[2 * t + 1, g]), (B[2 * t + 8, g], B[2 * t + 9, g]) = b_frag[lane][0], b_frag[lane][1]
        C[g, 2 * t], C[g, 2 * t + 1], C[g + 8, 2 * t], C[g + 8, 2 * t + 1] = c_frag[lane]
    D = (A.astype(np.float64) @ B.astype(np.float64)).astype(np.float32) + C
    out = np.zeros((32, 4), np.float32)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        out[lane] = D[g, 2 * t], D[g, 2 * t + 1], D[g + 8, 2 * t], D[g + 8, 2 * t + 1]
    return out


def column_pass_model(F, w, radius, n_valid):
    """One 16-column tile of one item: F[f][col] filtered rows (fp32), w[d] weights (fp32)."""
    R, N = radius, nch(radius)
    n_chunks = steps(n_valid + 2 * R) // 4 * (RING // CHUNK)
    F = np.vstack([F, np.zeros((n_chunks * CHUNK - F.shape[0], 16), np.float32)])
    f16 = lambda x: np.float32(np.float16(x))
    bh = np.zeros((N, 32, 2), np.float32)        # tf32 operand: hi(B_j[t][g]), hi(B_j[t+4][g])
    bc = np.zeros((N, 32, 2, 2), np.float32)     # fp16 operand: (w, w) pair and (w_lo, w_lo) pair
    for j in range(N):
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            ws = []
            for h in range(2):
                d = abs(8 * j + t + 4 * h - g - R)
                ws.append(np.float32(w[d]) if d <= R else np.float32(0))
            bh[j, lane] = [tf32_hi(x) for x in ws]
            bc[j, lane, 0] = [f16(x) for x in ws]
            bc[j, lane, 1] = [f16(np.float32(x) - tf32_hi(x)) for x in ws]
    acc = np.zeros((N - 1, 32, 4), np.float32)
    got = np.full((n_valid, 16), np.nan, np.float32)
    for c in range(n_chunks):
        ah = np.zeros((32, 4), np.float32)
        ac = np.zeros((32, 4, 2), np.float32)
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            a = np.array([F[8 * c + t, 2 * g], F[8 * c + t, 2 * g + 1], F[8 * c + t + 4, 2 * g], F[8 * c + t + 4, 2 * g + 1]],
                         np.float32)
            hi = tf32_hi(a)
            lo = a - hi
            ah[lane] = hi
            ac[lane] = [(f16(lo[0]), f16(lo[2])), (f16(lo[1]), f16(lo[3])), (f16(hi[0]), f16(hi[2])), (f16(hi[1]), f16(hi[3]))]
        nxt = np.zeros((N, 32, 4), np.float32)
        for j in range(N - 1, -1, -1):
            base = np.zeros((32, 4), np.float32) if j == 0 else acc[j - 1]
            nxt[j] = mma(ac, bc[j], base, 16)
            nxt[j] = mma(ah, bh[j], nxt[j], 8)
        acc[:] = nxt[:N - 1]
        ob = 8 * (c - (N - 1))
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for e, row in ((0, ob + 2 * t), (1, ob + 2 * t + 1)):
                if 0 <= row < n_valid:
                    got[row, 2 * g], got[row, 2 * g + 1] = nxt[N - 1, lane, e], nxt[N - 1, lane, e + 2]
    return got


@pytest.mark.parametrize("radius,n_valid", [(11, 37), (11, 8), (10, 29), (9, 21), (7, 40), (5, 13), (3, 1)])
def test_split_precision_banded_products_match_the_fp64_filter(radius, n_valid):
    rng = np.random.default_rng(100 + radius)
    sigma = radius / 5.5
    x = np.arange(-radius, radius + 1)
    full = np.exp(-0.5 * (x / sigma) ** 2)
    full /= full.sum()
    w = full[radius:].astype(np.float32)                       # w[d], like GaussParams<float>
    F = rng.random((n_valid + 2 * radius, 16), dtype=np.float32)
    got = column_pass_model(F, w, radius, n_valid)
    want = np.zeros((n_valid, 16))
    for o in range(n_valid):
        for f in range(o, o + 2 * radius + 1):
            want[o] += float(w[abs(f - o - radius)]) * F[f].astype(np.float64)
    assert not np.isnan(got).any()
    # dropped lo.lo term and the rounding of the lo parts: each below 2^-21 of the result
    assert np.abs(got - want).max() <= 1e-6


def test_split_is_exact_and_corrections_fit_fp16():
    rng = np.random.default_rng(7)
    a = rng.random(10000, dtype=np.float32)
    hi = tf32_hi(a)
    lo = a - hi
    assert np.array_equal((hi.astype(np.float64) + lo.astype(np.float64)).astype(np.float32), a)
    assert np.all(np.abs(lo) <= np.abs(a) * 2.0 ** -10)
    assert np.array_equal(np.float16(hi).astype(np.float32), hi)       # 11 significant bits: exact in fp16
    assert np.all(np.abs(np.float16(lo).astype(np.float32) - lo) <= np.maximum(np.abs(lo) * 2.0 ** -11, 2.0 ** -25))


# ------------------------------------------------------------------ hand-off protocol
def simulate_item(n_rows, producer_lead):
    """Replay one work item.  Returns the number of chunks consumed.  `producer_lead` bounds how far
    the producers run ahead when they could (0 = as late as the consumer allows, large = as early
    as the ring allows), to exercise both extremes of the schedule."""
    n_steps = steps(n_rows)
    n_chunks = n_steps // 4 * (RING // CHUNK)
    assert n_steps % GROUPS == 0 and n_chunks * CHUNK == n_steps * ROWS_PER_GROUP >= n_rows + 7
    produced = 0            # groups written (rows [12 g, 12 g + 12) of the item)
    released = 0            # groups handed back by the consumer
    waited = 0
    owner = [None] * RING   # item row currently held by each ring row
    for c in range(n_chunks):
        need = (8 * c + 7) // ROWS_PER_GROUP + 1
        # the producers write group g only once its ring slot was released (they block on empty[g % 4]
        # otherwise): g - released < GROUPS
        target = min(n_steps, need + producer_lead, released + GROUPS)
        assert target >= need, "deadlock: the consumer waits for a group the producers may not write yet"
        while produced < target:
            for r in range(ROWS_PER_GROUP):
                row = produced * ROWS_PER_GROUP + r
                owner[row % RING] = row
            produced += 1
        while waited < need:
            assert waited < produced
            waited += 1
        for r in range(8 * c, 8 * c + 8):                       # the chunk's rows are the item's rows
            assert owner[r % RING] == r, "chunk read before / after its rows were in the ring"
        done = (8 * (c + 1)) // ROWS_PER_GROUP
        assert done <= waited
        released = max(released, done)
    assert released == waited == n_steps                         # ring empty, barriers in phase
    # every output row's support is inside chunks the schedule delivers before the item ends
    return n_chunks


@pytest.mark.parametrize("lead", [0, 1, 2, 100])
def test_ring_protocol_never_reads_early_or_overwrites_late(lead):
    for radius in (3, 5, 7, 9, 10, 11):
        for n_valid in list(range(1, 60)) + [135, 540, 1080, 2160]:
            n_rows = n_valid + 2 * radius
            n_chunks = simulate_item(n_rows, lead)
            last_block = (n_valid - 1) // 8
            assert last_block + nch(radius) - 1 < n_chunks, "an output block would complete after the item ends"


def test_producer_can_always_make_progress():
    """Deadlock freedom: at the moment the consumer waits for group `need - 1`, the producers are
    allowed to write it (its ring slot has been released)."""
    for n_rows in range(1, 400):
        n_steps = steps(n_rows)
        released = 0
        for c in range(n_steps // 4 * (RING // CHUNK)):
            need = (8 * c + 7) // ROWS_PER_GROUP + 1
            assert (need - 1) - released < GROUPS
            released = (8 * (c + 1)) // ROWS_PER_GROUP


# ------------------------------------------------------------------ skewed transpose layout
def test_skewed_transpose_layout_is_conflict_free_and_tma_aligned():
    """transpose_tma64_kernel<K> (kernels/geometry.cuh): row y of a tile lands at
    y * 96 + 4 * ((y >> S) & 7) words.  Every warp-wide LDS.32 (K = 1) and every half-warp of an
    LDS.64 (K = 2) of the gather must hit distinct banks, rows must stay inside the pitch, and every
    bulk-copy destination must be 16-byte aligned."""
    pitch = 96
    for K, S in ((1, 2), (2, 1)):
        tx = 64 // K
        for y in range(64):
            start = y * pitch + 4 * ((y >> S) & 7)
            assert (start * 4) % 16 == 0 and 4 * ((y >> S) & 7) + 64 <= pitch
        for tw in (tx, tx - 4, 8, 4):                      # full and partial tile widths (multiples of 4 / K ... even)
            vpr = 64 * K // 4
            for i0 in range(0, 8 * tw * ((vpr + 7) // 8), 32):
                lanes = []
                for i in range(i0, i0 + 32):
                    q_lo, t2 = i & 7, i >> 3
                    q_hi, r = divmod(t2, tw)
                    q = 8 * q_hi + q_lo
                    if q < vpr:
                        lanes.append(((q << S) * pitch + 4 * q_lo + r * K, q, r))
                if K == 1:
                    for e in range(4):                     # the four LDS.32 of a thread: rows 4q + e
                        banks = [(a + e * pitch) % 32 for a, _, _ in lanes]
                        assert len(set(banks)) == len(banks), (K, tw, i0, e)
                else:
                    for half in (lanes[:16], lanes[16:]):  # LDS.64: 16 lanes x 2 banks per wavefront
                        for e in range(2):
                            banks = [b for a, _, _ in half for b in ((a + e * pitch) % 32, (a + e * pitch + 1) % 32)]
                            assert len(set(banks)) == len(banks), (K, tw, i0, e)
                            assert all(a % 2 == 0 for a, _, _ in half)


# ------------------------------------------------------------------ work items of a streaming-Gaussian launch
def _plan(columns, height, radius, sms, mma=1):
    import ctypes

    from millipyde_b200 import capi
    out = (ctypes.c_int * 4)()
    capi.lib().mpimg_gauss_stream_plan(columns, height, radius, sms, mma, out)
    return tuple(out)


def _gs_item(plan, n_strips, height, item):
    """kernels/gaussian_stream.cuh: gs_item."""
    main_items, tail_cols, chunk_rows, _ = plan
    if item < main_items:
        col, y0, y1 = item, 0, height
    else:
        chunk, j = divmod(item - main_items, tail_cols)
        col, y0 = main_items + j, chunk * chunk_rows
        y1 = min(height, y0 + chunk_rows)
    return col // n_strips, col % n_strips, y0, y1


def _row_steps(n_out, radius, mma=1):
    n_rows = n_out + 2 * radius
    return steps(n_rows) * ROWS_PER_GROUP if mma else (n_rows + 9) // 10 * 10


@pytest.mark.parametrize("n_images,height,n_strips,radius,sms", [
    (256, 2160, 18, 11, 148),     # the headline launch: 31 whole waves + 20 columns
    (1, 2160, 18, 11, 148),       # one 4K image: all tail
    (64, 2160, 18, 11, 148), (16, 1080, 9, 11, 148), (1024, 1080, 9, 11, 148), (3, 770, 1, 3, 148),
    (37, 45, 2, 11, 148), (148, 300, 1, 7, 148), (149, 300, 1, 7, 148), (5, 7, 3, 11, 4), (1, 1, 1, 3, 148),
])
def test_stream_items_cover_every_row_of_every_column_once(n_images, height, n_strips, radius, sms):
    columns = n_images * n_strips
    plan = _plan(columns, height, radius, sms)
    main_items, tail_cols, chunk_rows, n_chunks = plan
    n_items = main_items + tail_cols * n_chunks
    assert main_items % sms == 0 and main_items <= columns
    assert n_items == main_items or main_items + tail_cols == columns
    covered = np.zeros((columns, height), np.int32)
    for item in range(n_items):
        img, strip, y0, y1 = _gs_item(plan, n_strips, height, item)
        assert 0 <= img < n_images and 0 <= strip < n_strips and 0 <= y0 < y1 <= height   # every item owns a row
        covered[img * n_strips + strip, y0:y1] += 1
    assert (covered == 1).all()


def test_stream_items_tail_is_a_short_wave():
    """256 x 4K RGB on 148 SMs (bench.py's launch): the 20 columns behind the 31 whole waves go
    through as one wave of short items, and the launch's makespan drops below 32 whole waves."""
    plan = _plan(256 * 18, 2160, 11, 148)
    main_items, tail_cols, chunk_rows, n_chunks = plan
    assert (main_items, tail_cols) == (31 * 148, 20)
    assert tail_cols * n_chunks <= 148
    per_cta = np.zeros(148, np.int64)
    for item in range(main_items + tail_cols * n_chunks):
        _, _, y0, y1 = _gs_item(plan, 18, 2160, item)
        per_cta[item % 148] += _row_steps(y1 - y0, 11)
    whole_waves = 32 * _row_steps(2160, 11)
    assert per_cta.max() < 0.98 * whole_waves
    # a single image is cut as before: 18 columns x 8 chunks = 144 items in one wave
    assert _plan(18, 2160, 11, 148) == (0, 18, 270, 8)


# ------------------------------------------------------------------ staged boxes of the rotates
def _rot_params(w, h, angle_deg, centre):
    th = np.deg2rad(angle_deg)
    return np.cos(th), np.sin(th), centre[0], centre[1]


@pytest.mark.parametrize("tile_h,box,margin_lo,margin_hi,centre_rule", [
    (32, 50, 1, 2, "oracle"),     # gather_f32_kernel, 32 x 32 tiles: bx0 = floor(xmin) - 1, bw = ceil(xmax) + 2 - bx0
    (64, 76, 1, 2, "oracle"),     # gather_f32_kernel, single-channel 32 x 64 tiles
    (64, 78, 2, 3, "reference"),  # rotate_box_kernel<1>: bx0 = floor(xmin) - 2, bw = ceil(xmax) + 3 - bx0
    (32, 52, 2, 3, "reference"),  # rotate_box_kernel<2>
    (32, 52, 2, 3, "oracle"),     # rotate_box_kernel<2, bilinear>
])
def test_staged_box_holds_every_corner_of_every_sample(tile_h, box, margin_lo, margin_hi, centre_rule):
    """kernels/geometry.cuh: the box a block stages must contain both corners (floor and floor + 1) of
    every sample of its tile in both directions, and never exceed the compiled box edge -- for any
    angle, tile position and ragged image edge.  (The kernels read a coordinate outside the box from
    global memory or clamp nothing, so this is what their speed AND the gather's correctness rest on.)"""
    rng = np.random.default_rng(tile_h * 100 + box)
    for _ in range(400):
        w, h = int(rng.integers(33, 700)), int(rng.integers(33, 700))
        angle = float(rng.uniform(-180, 180)) if rng.random() < 0.8 else float(rng.choice([0, 30, 45, 90, 135, 180, -45]))
        cx, cy = ((w / 2 - 0.5, h / 2 - 0.5) if centre_rule == "oracle" else (w / 2, h / 2))
        c, s, cx, cy = _rot_params(w, h, angle, (cx, cy))
        ox0 = 32 * int(rng.integers(0, (w + 31) // 32))
        oy0 = tile_h * int(rng.integers(0, (h + tile_h - 1) // tile_h))
        ox1, oy1 = min(ox0 + 32, w) - 1, min(oy0 + tile_h, h) - 1
        xs, ys = np.meshgrid(np.arange(ox0, ox1 + 1, dtype=np.float64), np.arange(oy0, oy1 + 1, dtype=np.float64))
        sx = c * (xs - cx) - s * (ys - cy) + cx
        sy = s * (xs - cx) + c * (ys - cy) + cy
        # the kernels' box: extremes of the affine map over the tile (analytic half-extents == corner extremes)
        xmin, xmax, ymin, ymax = sx.min(), sx.max(), sy.min(), sy.max()
        bx0, by0 = int(np.floor(xmin)) - margin_lo, int(np.floor(ymin)) - margin_lo
        bw_need, bh_need = int(np.ceil(xmax)) + margin_hi - bx0, int(np.ceil(ymax)) + margin_hi - by0
        assert bw_need <= box and bh_need <= box, (w, h, angle, bw_need, bh_need)
        if centre_rule == "reference":       # nearest rule: truncation toward zero
            ix, iy = np.trunc(sx).astype(int), np.trunc(sy).astype(int)
            assert ix.min() >= bx0 and ix.max() < bx0 + bw_need and iy.min() >= by0 and iy.max() < by0 + bh_need
        ix, iy = np.floor(sx).astype(int), np.floor(sy).astype(int)   # bilinear: corners floor and floor + 1
        assert ix.min() >= bx0 and ix.max() + 1 < bx0 + bw_need
        assert iy.min() >= by0 and iy.max() + 1 < by0 + bh_need


# ------------------------------------------------------------------ TMA alignment of the staged boxes / tiles
def test_bulk_copies_of_the_staged_boxes_are_16_byte_aligned():
    """cp.async.bulk needs 16-byte-aligned source, destination and size.  The box kernels
    (gather_f32_kernel, rotate_box_kernel: C or K 32-bit words per pixel) copy, per box row, the words
    [c_lo, c_hi) of the enclosing aligned span; the launchers only take this path when image rows are
    whole vectors ((W * K) % 4 == 0) and the base is aligned.  For any box origin, width and pitch
    (a multiple of 4 words) every copy must then be aligned on both sides -- a violation is a fault on
    the device, not a wrong pixel."""
    rng = np.random.default_rng(99)
    for _ in range(4000):
        K = int(rng.choice([1, 2, 3, 4]))
        W = int(rng.integers(8, 900))
        while (W * K) % 4:
            W += 1
        H = int(rng.integers(8, 900))
        bx0, by0 = int(rng.integers(-80, W + 10)), int(rng.integers(-80, H + 10))
        bw, bh = int(rng.integers(1, 79)), int(rng.integers(1, 79))
        pitch = 4 * ((bw * K + 6) // 4 + 1 + int(rng.integers(0, 6)))
        row_len = W * K
        shift = (bx0 * K) & 3                      # Python's & on negatives is two's complement too
        want = (shift + bw * K + 3) & ~3
        col0 = bx0 * K - shift
        c_lo = -col0 if col0 < 0 else 0
        c_hi = min(want, row_len - col0)
        r_lo = -by0 if by0 < 0 else 0
        r_hi = min(bh, H - by0)
        if not (c_hi > c_lo and r_hi > r_lo):
            continue
        assert col0 % 4 == 0 and c_lo % 4 == 0 and c_hi % 4 == 0 and want % 4 == 0
        assert 0 <= shift < 4 and shift + bw * K <= want <= pitch + 4
        for r in (r_lo, r_hi - 1):
            src_word = (by0 + r) * row_len + col0 + c_lo
            dst_word = r * pitch + c_lo
            assert src_word % 4 == 0 and dst_word % 4 == 0 and (c_hi - c_lo) % 4 == 0
            assert 0 <= (by0 + r) < H and 0 <= col0 + c_lo and col0 + c_hi <= row_len      # inside the image row


def test_bulk_copies_of_the_f64_gaussian_tile_are_16_byte_aligned():
    """gauss_f64_kernel's TMA staging (even widths only): per staged row the in-image columns
    [c_lo, c_hi) of the tile [x0 - R, x0 + 32 + R), in doubles."""
    for R in (8, 16):
        in_w, in_h, th = 32 + 2 * R, 128, 128 - 2 * R
        pitch_in = in_w + 2
        for W in range(2, 400, 2):
            for H in (1, 37, th, th + 1, 300):
                for bx in range((W + 31) // 32):
                    for by in range((H + th - 1) // th):
                        gx_lo, gy_lo = 32 * bx - R, th * by - R
                        c_lo, c_hi = (-gx_lo if gx_lo < 0 else 0), min(in_w, W - gx_lo)
                        r_lo, r_hi = (-gy_lo if gy_lo < 0 else 0), min(in_h, H - gy_lo)
                        assert c_hi > c_lo and r_hi > r_lo            # a launched tile always owns pixels
                        assert (c_hi - c_lo) % 2 == 0 and c_lo % 2 == 0 and (gx_lo + c_lo) % 2 == 0
                        assert (r_lo * pitch_in + c_lo) % 2 == 0 and pitch_in % 2 == 0
