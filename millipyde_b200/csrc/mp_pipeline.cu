// Operation-chain executor: coin flips -> grouping -> fusion pass -> batched
// launches -> sharding over devices -> peer hand-off between connected pipelines.
//
// Replaces PyGPUPipeline_run / gpupipeline_run_sequence / gpupipeline_send_input
// (src/gpupipeline.c:234-403) and the serial Generator loop
// (src/gpugenerator.c:203-281).  See include/mp_pipeline.h for the contract.
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <unordered_set>
#include <vector>

#include "mp_devices.h"
#include "mp_image.h"
#include "mp_internal.h"
#include "mp_objects.h"
#include "mp_ops_internal.h"
#include "mp_pipeline.h"
#include "mp_internal_rng.h"
#include "mp_rng.h"
#include "kernels/records.cuh"

using namespace mpk;

namespace {

enum OpKind {
    OP_GREY, OP_TRANSPOSE, OP_GAUSSIAN, OP_FLIPLR, OP_ROTATE, OP_BRIGHTNESS, OP_GAMMA, OP_COLORIZE,
    OP_ELEMENTWISE,  // mpimg_elementwise: a ufunc-exact pointwise op (fuses like the other pointwise ops on fp32)
    OP_RANDOM,   // one of the mpimg_random_* operators: parameters are drawn per image at run time
    OP_FOREIGN,  // an MPFunc that is not one of ours: called as-is, one image at a time
    OP_NULL,     // func == NULL: skipped, as src/gpupipeline.c:393-396
};

struct Stage {
    OpKind kind;
    MPFunc func;
    void *args;          // owned copy for our operators, caller's pointer for foreign ones
    double a[5];         // the leading doubles of the args block (ElementwiseArgs is the longest)
    double probability;  // <= 0: always (the reference tests `probability > 0`, :380)
    // a realised random_* stage remembers where its parameters came from: a[i] is the draw
    // keyed_range(run, image, rstage, slot i, lo[i], hi[i]) (mp_rng.h); kFixedStage = not drawn
    uint32_t rstage = kFixedStage;
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
};

std::atomic<int> g_device_draws{1};   // fill per-image records on the device when a launch's images share the template

std::atomic<int> g_fusion{1};

bool is_pointwise(OpKind k);
// Can this stage ride in a fused pointwise program of an image with `channels` channels?  (A per-channel
// np.multiply needs exactly the three factors a PwOp carries.)
bool is_pw(const Stage &s, int channels)
{
    if (!is_pointwise(s.kind)) return false;
    if (s.kind == OP_ELEMENTWISE && (int)s.a[0] == MP_EW_MUL && s.a[4] != 0 && channels != 3) return false;
    return true;
}

OpKind classify(MPFunc f, size_t *arg_bytes)
{
    *arg_bytes = 0;
    if (!f) return OP_NULL;
    if (f == mpimg_color_to_greyscale) return OP_GREY;
    if (f == mpimg_transpose) return OP_TRANSPOSE;
    if (f == mpimg_fliplr) return OP_FLIPLR;
    if (f == mpimg_gaussian) { *arg_bytes = sizeof(GaussianArgs); return OP_GAUSSIAN; }
    if (f == mpimg_rotate) { *arg_bytes = sizeof(RotateArgs); return OP_ROTATE; }
    if (f == mpimg_brightness) { *arg_bytes = sizeof(BrightnessArgs); return OP_BRIGHTNESS; }
    if (f == mpimg_adjust_gamma) { *arg_bytes = sizeof(GammaArgs); return OP_GAMMA; }
    if (f == mpimg_colorize) { *arg_bytes = sizeof(ColorizeArgs); return OP_COLORIZE; }
    if (f == mpimg_elementwise) { *arg_bytes = sizeof(ElementwiseArgs); return OP_ELEMENTWISE; }
    if (f == mpimg_random_rotate || f == mpimg_random_gaussian || f == mpimg_random_brightness) {
        *arg_bytes = sizeof(RandomRangeArgs);
        return OP_RANDOM;
    }
    if (f == mpimg_random_adjust_gamma) { *arg_bytes = sizeof(RandomGammaArgs); return OP_RANDOM; }
    if (f == mpimg_random_colorize) { *arg_bytes = sizeof(RandomColorizeArgs); return OP_RANDOM; }
    return OP_FOREIGN;
}

bool is_pointwise(OpKind k) { return k == OP_BRIGHTNESS || k == OP_GAMMA || k == OP_COLORIZE || k == OP_ELEMENTWISE; }

// Per-device page-locked arenas for pointer tables.  A shard's worker owns one arena of the device's
// ring while it enqueues (under the device's `enqueue` mutex, which also keeps the launches of two
// shards from interleaving on the stream) and lets go of the mutex before it waits for the device,
// so the next shard's host work overlaps this shard's kernels.  `done` marks the end of the last
// shard that used the arena: it is what that shard's worker waits on, and what guards reuse.
struct Arena {
    char *base = nullptr;
    size_t cap = 0, used = 0;
    cudaEvent_t done = nullptr;
    bool in_use = false;   // `done` has been recorded and not yet waited for by a new owner
    void *take(size_t bytes)
    {
        bytes = (bytes + 63) & ~(size_t)63;
        if (used + bytes > cap) return nullptr;
        void *p = base + used;
        used += bytes;
        return p;
    }
};
constexpr int kArenaRing = 2;
struct DeviceArenas {
    std::mutex enqueue;
    Arena ring[kArenaRing];
    unsigned next = 0;
};
DeviceArenas g_arenas[64];
thread_local Arena *g_arena = nullptr;   // the arena of the shard this worker thread is enqueueing

}  // namespace

struct mp_pipeline {
    std::vector<Stage> stages;
    int device = DEVICE_LOC_NO_AFFINITY;
    mp_pipeline *receiver = nullptr;
    std::atomic<unsigned long long> launches{0};
    std::atomic<int> segments{0};
    // random source of the run in flight: every draw is keyed_*(run_key, index_base + position of the
    // image in the submitted array, stage, slot) -- independent of devices, shards and threads
    uint64_t run_key = 0;
    uint64_t index_base = 0;
    bool key_held = false;   // mppipe_hold_run_key: one key for every run (a Generator's whole stream)
    std::atomic<int> status{MILLIPYDE_SUCCESS};
    // devices touched by the run in flight (for mppipe_wait)
    std::mutex mux;
    std::vector<int> in_flight;
    bool cycled = false;
    // shards of this pipeline submitted and not yet finished; mppipe_wait of a pipeline with a fixed
    // device waits for this to reach zero instead of draining the whole device, so two pipelines can
    // keep one device busy back to back (submit B, wait A, submit A, wait B, ...)
    int pending = 0;
    bool soft_wait = false;
    std::condition_variable cv;
};

namespace {

void note_status(mp_pipeline *p, MPStatus st)
{
    if (st != MILLIPYDE_SUCCESS) {
        int expected = MILLIPYDE_SUCCESS;
        p->status.compare_exchange_strong(expected, (int)st);
    }
}

// ------------------------------------------------------------------ fusion pass
struct Segment {
    enum Kind { SINGLE, PW_F32, GREY_F32, PW_RGBA8, GATHER_F32, GAUSS_F32 } kind;
    const Stage *single = nullptr;  // SINGLE; GAUSS_F32: the gaussian stage (sigma = single->a[0])
    PwProgram pre = {}, post = {};  // PW_F32 uses `pre`; GREY_F32, GATHER_F32 and GAUSS_F32 use both
    U8Program u8 = {};
    // GATHER_F32: flips before / after the (optional) rotate, and its angle
    bool flip_pre = false, flip_post = false, has_rotate = false;
    double angle = 0;
    // where the values above came from (fixed arguments or keyed draws): what the device-side record
    // fill evaluates when every image of a launch shares it
    PwTemplate pre_t = {}, post_t = {};
    // GAUSS_F32: what the streaming kernel is handed.  k_pre = pre + the per-channel factors of the
    // ops behind the blur (the blur is linear, so they run in front of it); k_post = empty or ONE op
    // read as out = min(max(v + a, b), c), what the add / clamp parts of those ops compose to.
    PwProgram k_pre = {}, k_post = {};
    PwTemplate k_pre_t = {}, k_post_t = {};
    bool tmpl_ok = true;   // the templates reproduce the values (false: a drawn parameter went through arithmetic)
    uint32_t angle_stage = kFixedStage;
    double angle_lo = 0, angle_hi = 0;
};

PwOp to_pw(const Stage &s)
{
    switch (s.kind) {
        case OP_BRIGHTNESS: return PwOp{PW_BRIGHTNESS, (float)s.a[0], 0.f, 0.f};
        case OP_GAMMA: return PwOp{PW_GAMMA, (float)s.a[0], (float)s.a[1], 0.f};
        case OP_ELEMENTWISE: {   // a = {kind, a, b, c, per_channel}; uniform factors are replicated
            const bool per_channel = s.a[4] != 0;
            const float m = (float)s.a[1];
            if ((int)s.a[0] == MP_EW_MUL && !per_channel) return PwOp{PW_EW_MUL, m, m, m};
            return PwOp{(int)s.a[0], m, (float)s.a[2], (float)s.a[3]};
        }
        default: return PwOp{PW_COLORIZE, (float)s.a[0], (float)s.a[1], (float)s.a[2]};
    }
}

PwTemplateOp to_pw_t(const Stage &s)
{
    const PwOp v = to_pw(s);
    PwTemplateOp t = {};
    t.kind = v.kind;
    t.stage = s.rstage;
    if (s.rstage == kFixedStage) {
        t.lo[0] = t.hi[0] = v.a;
        t.lo[1] = t.hi[1] = v.b;
        t.lo[2] = t.hi[2] = v.c;
    } else {   // brightness: (delta); gamma: (gamma, gain); colorize: (r, g, b) -- the PwOp's a, b, c in order
        const int np = s.kind == OP_BRIGHTNESS ? 1 : (s.kind == OP_GAMMA ? 2 : 3);
        for (int i = 0; i < np; ++i) {
            t.lo[i] = s.lo[i];
            t.hi[i] = s.hi[i];
        }
    }
    return t;
}

// append a pointwise stage to a segment's program (value form and template form stay in step)
void push_pw(PwProgram &prog, PwTemplate &tmpl, const Stage &s)
{
    tmpl.ops[tmpl.n++] = to_pw_t(s);
    prog.ops[prog.n++] = to_pw(s);
}

U8Op to_u8(const Stage &s)
{
    switch (s.kind) {
        case OP_BRIGHTNESS: return mp::u8_brightness_op(s.a[0]);
        case OP_GAMMA: return U8Op{PW_GAMMA, 0, s.a[0], s.a[1], 0};
        default: return U8Op{PW_COLORIZE, 0, s.a[0], s.a[1], s.a[2]};
    }
}

// ---- pointwise ops behind a Gaussian, folded -------------------------------------------------
// Every op the fold accepts is g(y) = clamp(a_ch y + b, l, h) with a_ch >= 0; compositions of such maps
// are again of that form, f(v) = clamp(A_ch v + B, L, H).  The kernel applies v + B and the clamp to the
// finished rows (three instructions on the accumulator fragments) and A_ch -- by linearity of the blur --
// to the samples in front of it.  An op that would make B, L or H differ between channels, or that is
// not affine (gamma, power), ends the fold: it and what follows run as a segment of their own.
struct PostFold {
    double A[3] = {1, 1, 1};
    double B = 0, L = -HUGE_VAL, H = HUGE_VAL;
    int n = 0;              // ops absorbed
    bool drawn_other = false;   // a drawn parameter other than "one brightness" took part
    const Stage *only_brightness = nullptr;
};

bool fold_post(PostFold &f, const Stage &t, int channels)
{
    double a[3] = {1, 1, 1}, b = 0, l = -HUGE_VAL, h = HUGE_VAL;
    bool per_channel = false;
    switch (t.kind) {
        case OP_BRIGHTNESS:
            if (channels == 4) return false;   // skips alpha: not the same map on every channel
            b = t.a[0]; l = 0; h = 1;
            break;
        case OP_COLORIZE:
            if (channels == 1) { ++f.n; return true; }   // no-op on grey (:647-651)
            if (channels != 3) return false;
            a[0] = t.a[0]; a[1] = t.a[1]; a[2] = t.a[2]; h = 1;
            per_channel = true;
            break;
        case OP_ELEMENTWISE: {
            const int kind = (int)t.a[0];
            if (kind == MP_EW_ADD) b = t.a[1];
            else if (kind == MP_EW_CLIP) { l = t.a[2]; h = t.a[3]; if (!(l <= h)) return false; }
            else if (kind == MP_EW_MUL) {
                if (t.a[4] != 0) { a[0] = t.a[1]; a[1] = t.a[2]; a[2] = t.a[3]; per_channel = true; }
                else a[0] = a[1] = a[2] = t.a[1];
            } else return false;
            break;
        }
        default: return false;
    }
    for (int c = 0; c < 3; ++c)
        if (!(a[c] > 0) || !std::isfinite(a[c])) return false;   // clamp bounds would swap (or collapse)
    if (per_channel && (a[0] != a[1] || a[1] != a[2])) {
        // channel-dependent slope: B, L, H stay common to the channels only from this state
        if (f.B != 0 || !(f.L == -HUGE_VAL || f.L == 0) || f.H != HUGE_VAL) return false;
        if (channels != 3) return false;
    }
    auto clampd = [](double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); };
    const double u = a[0];   // the slope applied to B, L, H (uniform, or they are 0 / infinite)
    for (int c = 0; c < 3; ++c) f.A[c] *= a[c];
    f.B = u * f.B + b;
    f.L = clampd(f.L == -HUGE_VAL ? -HUGE_VAL : u * f.L + b, l, h);
    f.H = clampd(f.H == HUGE_VAL ? HUGE_VAL : u * f.H + b, l, h);
    if (t.rstage != kFixedStage) {
        if (t.kind == OP_BRIGHTNESS && f.n == 0) f.only_brightness = &t;
        else f.drawn_other = true;
    } else if (f.only_brightness) {
        f.drawn_other = true;   // a drawn delta followed by more ops: its bounds went through arithmetic
    }
    ++f.n;
    return true;
}

// Compile the surviving stages of one group into segments.  `fam`/`channels`
// track the layout as ops change it.  Legality rules (DESIGN.md "Fusion"):
//  * consecutive pointwise ops compose exactly -> one kernel (fp32: op program;
//    RGBA8: composed byte tables, bit-identical to running them one by one);
//  * rgb2grey absorbs the pointwise ops before it (per colour channel) and after
//    it (on the grey value) -> one kernel reading C channels, writing one;
//  * everything else is a segment of its own.
std::vector<Segment> compile(const std::vector<const Stage *> &ops, mp::Family fam, int channels)
{
    std::vector<Segment> out;
    const bool fuse = g_fusion.load() != 0;
    size_t i = 0;
    while (i < ops.size()) {
        const Stage *s = ops[i];
        // fliplr / rotate (+ the pointwise ops around them) -> one gather pass
        if (fuse && fam == mp::FAM_F32 && (s->kind == OP_FLIPLR || s->kind == OP_ROTATE || is_pw(*s, channels))) {
            Segment seg;
            seg.kind = Segment::GATHER_F32;
            size_t j = i;
            bool geometric = false;
            while (j < ops.size()) {
                const Stage *t = ops[j];
                if (t->kind == OP_FLIPLR) {
                    (seg.has_rotate ? seg.flip_post : seg.flip_pre) ^= true;
                    geometric = true;
                } else if (t->kind == OP_ROTATE && !seg.has_rotate) {
                    seg.has_rotate = true;
                    seg.angle = t->a[0];
                    seg.angle_stage = t->rstage;
                    seg.angle_lo = t->rstage == kFixedStage ? t->a[0] : t->lo[0];
                    seg.angle_hi = t->rstage == kFixedStage ? t->a[0] : t->hi[0];
                    geometric = true;
                } else if (is_pw(*t, channels)) {
                    PwProgram &prog = seg.has_rotate ? seg.post : seg.pre;
                    if (prog.n == kMaxPw) break;
                    push_pw(prog, seg.has_rotate ? seg.post_t : seg.pre_t, *t);
                } else {
                    break;
                }
                ++j;
            }
            if (geometric && (j - i >= 2 || seg.has_rotate)) {  // a lone rotate is a gather too
                out.push_back(seg);
                i = j;
                continue;
            }
        }
        // a Gaussian absorbs the pointwise ops next to it: the ones before it are applied to every
        // sample as it lands in shared memory (once per sample, halo columns included; the zero
        // padding stays zero), the ones after it to the finished output rows before they are stored
        if (fuse && fam == mp::FAM_F32 && (is_pw(*s, channels) || s->kind == OP_GAUSSIAN)) {
            Segment seg;
            seg.kind = Segment::GAUSS_F32;
            size_t j = i;
            while (j < ops.size() && is_pw(*ops[j], channels) && seg.pre.n < kMaxPw) push_pw(seg.pre, seg.pre_t, *ops[j++]);
            if (j < ops.size() && ops[j]->kind == OP_GAUSSIAN && ops[j]->a[0] > 1e-15) {
                seg.single = ops[j++];
                PostFold fold;
                while (j < ops.size() && is_pw(*ops[j], channels) && seg.post.n < kMaxPw && seg.pre.n + 1 < kMaxPw &&
                       fold_post(fold, *ops[j], channels))
                    push_pw(seg.post, seg.post_t, *ops[j++]);
                if (seg.pre.n + seg.post.n > 0) {   // a bare Gaussian keeps its own (SINGLE) path
                    seg.k_pre = seg.pre;
                    seg.k_pre_t = seg.pre_t;
                    if (fold.A[0] != 1 || fold.A[1] != 1 || fold.A[2] != 1) {
                        const PwOp mul = {PW_EW_MUL, (float)fold.A[0], (float)fold.A[1], (float)fold.A[2]};
                        PwTemplateOp mt = {};
                        mt.kind = PW_EW_MUL;
                        mt.stage = kFixedStage;
                        for (int c = 0; c < 3; ++c) mt.lo[c] = mt.hi[c] = (&mul.a)[c];
                        seg.k_pre.ops[seg.k_pre.n++] = mul;
                        seg.k_pre_t.ops[seg.k_pre_t.n++] = mt;
                        for (int q = 0; q < seg.post.n; ++q)   // a drawn colorize / multiply: the product is not a plain draw
                            if (seg.post_t.ops[q].stage != kFixedStage && seg.post.ops[q].kind != PW_BRIGHTNESS) seg.tmpl_ok = false;
                    }
                    if (seg.post.n > 0 && (fold.B != 0 || fold.L != -HUGE_VAL || fold.H != HUGE_VAL)) {
                        seg.k_post.n = 1;
                        seg.k_post.ops[0] = PwOp{PW_EW_ADD, (float)fold.B, (float)fold.L, (float)fold.H};
                        PwTemplateOp pt = {};
                        pt.kind = PW_EW_ADD;
                        pt.stage = kFixedStage;
                        pt.lo[0] = pt.hi[0] = (float)fold.B;
                        pt.lo[1] = pt.hi[1] = (float)fold.L;
                        pt.lo[2] = pt.hi[2] = (float)fold.H;
                        if (fold.only_brightness && !fold.drawn_other) {   // out = clamp(v + delta, 0, 1), delta drawn: slot 0
                            pt.stage = fold.only_brightness->rstage;
                            pt.lo[0] = fold.only_brightness->lo[0];
                            pt.hi[0] = fold.only_brightness->hi[0];
                        } else if (fold.drawn_other) {
                            seg.tmpl_ok = false;
                        }
                        seg.k_post_t.n = 1;
                        seg.k_post_t.ops[0] = pt;
                    }
                    out.push_back(seg);
                    i = j;
                    continue;
                }
            }
        }
        if (fuse && fam == mp::FAM_F32 && (is_pw(*s, channels) || (s->kind == OP_GREY && channels >= 3))) {
            Segment seg;
            seg.kind = Segment::PW_F32;
            bool grey = false;
            size_t j = i;
            while (j < ops.size()) {
                const Stage *t = ops[j];
                if (is_pw(*t, grey ? 1 : channels)) {
                    PwProgram &prog = grey ? seg.post : seg.pre;
                    if (prog.n == kMaxPw) break;
                    if (!(grey && t->kind == OP_COLORIZE))  // colorize is a no-op on grey (:647-651)
                        push_pw(prog, grey ? seg.post_t : seg.pre_t, *t);
                    ++j;
                } else if (t->kind == OP_GREY && !grey && channels >= 3) {
                    grey = true;
                    seg.kind = Segment::GREY_F32;
                    ++j;
                } else {
                    break;
                }
            }
            if (j - i >= 1) {  // even a single op: the segment form is the one that batches
                if (grey) channels = 1;
                out.push_back(seg);
                i = j;
                continue;
            }
        }
        if (fuse && fam == mp::FAM_RGBA8 && is_pw(*s, channels) && s->kind != OP_ELEMENTWISE) {
            Segment seg;
            seg.kind = Segment::PW_RGBA8;
            size_t j = i;
            while (j < ops.size() && is_pw(*ops[j], channels) && ops[j]->kind != OP_ELEMENTWISE && seg.u8.n < kMaxPw)
                seg.u8.ops[seg.u8.n++] = to_u8(*ops[j++]);
            if (j - i >= 2) {
                out.push_back(seg);
                i = j;
                continue;
            }
        }
        Segment seg;
        seg.kind = Segment::SINGLE;
        seg.single = s;
        out.push_back(seg);
        if (s->kind == OP_GREY) {
            if (fam == mp::FAM_RGBA8 || fam == mp::FAM_U8_OTHER || fam == mp::FAM_F64_OTHER) fam = mp::FAM_F64;
            channels = 1;
        }
        ++i;
    }
    return out;
}

// Per-image realisation of the chain: coin flips (src/gpupipeline.c:380-387) and, for random_*
// stages, the parameter draws the reference makes inside the gpuimage method.  The result is a
// list of concrete stages (args point at the stage's own doubles).
//
// Every draw is a pure function of (p->run_key, image, stage index, slot) -- mp_rng.h -- so the
// stream an image sees does not depend on which device, shard or worker thread handles it, costs no
// syscall, and can be re-evaluated on the device (kernels/records.cuh).
void realize(const mp_pipeline *p, uint64_t image, std::vector<Stage> *out)
{
    const size_t ns = p->stages.size();
    const uint64_t key = p->run_key;
    out->clear();
    out->reserve(ns);  // args pointers below rely on no reallocation
    for (size_t k = 0; k < ns; ++k) {
        const Stage &st = p->stages[k];
        bool run = st.kind != OP_NULL;
        if (run && st.probability > 0 && mprng::keyed_u01(key, image, (uint32_t)k, mprng::kCoinSlot) > st.probability)
            run = false;   // runs iff rand <= p, src/gpuoperation.c:204
        if (!run) continue;
        if (st.kind != OP_RANDOM) {
            out->push_back(st);
            continue;
        }
        Stage c = st;
        c.probability = -1;
        c.rstage = (uint32_t)k;
        const double *r = (const double *)st.args;   // (min, max) pairs in parameter order
        int np = 1;
        if (st.func == mpimg_random_rotate) { c.kind = OP_ROTATE; c.func = mpimg_rotate; }
        else if (st.func == mpimg_random_gaussian) { c.kind = OP_GAUSSIAN; c.func = mpimg_gaussian; }
        else if (st.func == mpimg_random_brightness) { c.kind = OP_BRIGHTNESS; c.func = mpimg_brightness; }
        else if (st.func == mpimg_random_adjust_gamma) { c.kind = OP_GAMMA; c.func = mpimg_adjust_gamma; np = 2; }
        else { c.kind = OP_COLORIZE; c.func = mpimg_colorize; np = 3; }
        for (int i = 0; i < np; ++i) {
            c.lo[i] = r[2 * i];
            c.hi[i] = r[2 * i + 1];
            c.a[i] = mprng::keyed_range(key, image, (uint32_t)k, (uint32_t)i, c.lo[i], c.hi[i]);
        }
        out->push_back(c);
    }
    for (Stage &c : *out)
        if (c.kind != OP_FOREIGN && c.kind != OP_GREY && c.kind != OP_TRANSPOSE && c.kind != OP_FLIPLR) c.args = (void *)c.a;
}

// What decides the KERNEL a segment runs (never its parameter VALUES: those travel as per-image
// records).  Two images whose current segments have equal signatures and equal layouts share one
// launch.  The Gaussian adds the radius bucket its sigma falls in; with fusion off, pointwise and
// Gaussian values are part of the signature too, which restores the reference's one-image,
// one-op-at-a-time launches.
std::string signature(const Segment &g)
{
    char b[192];
    switch (g.kind) {
        case Segment::SINGLE: {
            const Stage &st = *g.single;
            if (st.kind == OP_GAUSSIAN) {
                int bucket = -1;
                if (st.a[0] > 1e-15 && g_fusion.load()) {
                    double w[kGaussMaxRadius + 1];
                    const int r = mp::oracle_weights(st.a[0], w, kGaussMaxRadius);
                    bucket = mp::gauss_stream_bucket(mp::effective_radius(w, r, mp::kGaussTailEps));
                }
                if (bucket > 0) snprintf(b, sizeof b, "S%d:G%d", (int)st.kind, bucket);
                else snprintf(b, sizeof b, "S%d:%.17g", (int)st.kind, st.a[0]);
            } else if (st.kind == OP_FOREIGN) {
                snprintf(b, sizeof b, "S%d:%p:%p", (int)st.kind, (void *)st.func, st.args);
            } else if (st.kind == OP_ROTATE && g_fusion.load()) {
                snprintf(b, sizeof b, "S%d", (int)st.kind);
            } else {
                snprintf(b, sizeof b, "S%d:%.17g,%.17g,%.17g,%.17g,%.17g", (int)st.kind, st.a[0], st.a[1], st.a[2], st.a[3],
                         st.a[4]);
            }
            break;
        }
        case Segment::GATHER_F32:
            snprintf(b, sizeof b, "T%d%d%d", (int)g.flip_pre, (int)g.flip_post, (int)g.has_rotate);
            break;
        case Segment::GAUSS_F32: {
            double w[kGaussMaxRadius + 1];
            const int r = mp::oracle_weights(g.single->a[0], w, kGaussMaxRadius);
            const int bucket = mp::gauss_stream_bucket(mp::effective_radius(w, r, mp::kGaussTailEps));
            if (bucket > 0) snprintf(b, sizeof b, "GP%d", bucket);
            else snprintf(b, sizeof b, "GP:%.17g", g.single->a[0]);   // no streaming kernel: runs op by op
            break;
        }
        case Segment::PW_RGBA8: {
            // composed byte tables are built per program: same program, same launch shape
            std::string k = "U";
            k.append((const char *)&g.u8, sizeof g.u8);
            return k;
        }
        default: snprintf(b, sizeof b, "P%d", (int)g.kind); break;
    }
    return b;
}

MPStatus run_gather(const std::vector<MPObjData *> &objs, const std::vector<const Segment *> &segs, const mp::Img &d,
                    int device, cudaStream_t s);

MPStatus run_segment_on(MPObjData *obj, const Segment &seg)
{
    switch (seg.kind) {
        case Segment::PW_F32: return mp::op_pointwise_f32(obj, seg.pre);
        case Segment::GREY_F32: return mp::op_grey_f32(obj, seg.pre, seg.post);
        case Segment::PW_RGBA8: return mp::op_pointwise_rgba8(obj, seg.u8);
        case Segment::GATHER_F32: {
            mp::Img d;
            if (!mp::describe(obj, &d) || d.fam != mp::FAM_F32) return MP_ERROR_UNSUPPORTED_LAYOUT;
            if (cudaSetDevice(obj->mem_loc) != cudaSuccess) return MP_ERROR_CUDA_RUNTIME;
            std::vector<MPObjData *> one(1, obj);
            std::vector<const Segment *> one_seg(1, &seg);
            return run_gather(one, one_seg, d, obj->mem_loc, mp::stream_of(obj));
        }
        case Segment::GAUSS_F32: {   // no fused form for this layout / shape / radius: op by op
            MPStatus st = seg.pre.n ? mp::op_pointwise_f32(obj, seg.pre) : MILLIPYDE_SUCCESS;
            if (st == MILLIPYDE_SUCCESS) st = seg.single->func(obj, seg.single->args);
            if (st == MILLIPYDE_SUCCESS && seg.post.n) st = mp::op_pointwise_f32(obj, seg.post);
            return st;
        }
        default: return seg.single->func(obj, seg.single->args);
    }
}

// Batched launch plumbing: fresh output buffers for every image of a same-shape group, device
// pointer tables (inputs then outputs) uploaded from the page-locked arena, one call of `launch`,
// then the inputs are retired in stream order.  *handled = false (and nothing changed) if the
// arena or the pool cannot serve the request.
//
// `records` (optional): per-image parameter records, uploaded behind the tables in the same copy;
// `launch` finds them at g_records (device address), valid for the duration of the call.
thread_local const void *g_records = nullptr;
// stream indices of the images of the group being launched (aligned with its objs) and the run's key:
// what the device-side record fill keys its draws with
thread_local const std::vector<uint64_t> *g_index = nullptr;
thread_local uint64_t g_run_key = 0;

// Views (mpobj_view_data) whose buffer still belongs to someone else, for the shard this worker
// thread is running.  The launch that first rewrites such an image must not free its input.
thread_local std::unordered_set<MPObjData *> *g_borrowed = nullptr;

// Device whose pool the outputs of the current batched launch come from: the shard's own device,
// or -- for the last segment of a chain that hands its images to a pipeline on another device -- the
// RECEIVER's, so the producer kernel itself moves the result over NVLink (no copy pass afterwards).
thread_local int g_out_device = -1;

// Retire the buffer an image held before a launch gave it a fresh one.
void release_input(int device, cudaStream_t s, MPObjData *o)
{
    if (g_borrowed && g_borrowed->erase(o)) return;  // not ours to free
    mp::pool_free(device, s, o->device_data);
}

// Give a still-borrowing view a buffer of its own (deep copy) before anything frees or mutates it.
MPStatus materialize(int device, cudaStream_t s, MPObjData *o)
{
    if (!g_borrowed || !g_borrowed->count(o)) return MILLIPYDE_SUCCESS;
    g_borrowed->erase(o);
    void *fresh = mp::pool_alloc(device, s, o->nbytes);
    if (!fresh) {
        o->device_data = NULL;
        return MP_ERROR_DEVICE_ALLOC;
    }
    cudaError_t e = cudaMemcpyAsync(fresh, o->device_data, o->nbytes, cudaMemcpyDeviceToDevice, s);
    o->device_data = fresh;
    if (e != cudaSuccess) {
        mp::record_cuda_error(e, "cudaMemcpyAsync(view)", __FILE__, __LINE__);
        return MP_ERROR_CUDA_RUNTIME;
    }
    return MILLIPYDE_SUCCESS;
}

// Device-side record fill (kernels/records.cuh): `records` then holds the launch's template(s) followed
// by the images' stream indices (8 bytes each, at offset index_offset); `fill` is called with the
// device addresses of both and of a fresh buffer of out_bytes, writes the per-image records there, and
// `launch` finds THAT buffer at g_records.
struct FillSpec {
    size_t index_offset = 0;
    size_t out_bytes = 0;
    std::function<void(const void *d_templates, const unsigned long long *d_index, void *d_out)> fill;
};

template <typename Launch>
MPStatus run_batched(const std::vector<MPObjData *> &objs, size_t out_bytes, int device, cudaStream_t s,
                     bool *handled, Launch launch, const void *records = nullptr, size_t record_bytes = 0,
                     const FillSpec *fill = nullptr)
{
    *handled = false;
    const size_t n = objs.size();
    const size_t tab_bytes = (2 * n * sizeof(void *) + 15) & ~(size_t)15;
    void **h_tab = g_arena ? (void **)g_arena->take(tab_bytes + record_bytes) : nullptr;
    if (!h_tab) return MILLIPYDE_SUCCESS;
    void *d_tab = mp::pool_alloc(device, s, tab_bytes + record_bytes);
    if (!d_tab) return MP_ERROR_DEVICE_ALLOC;
    void *d_filled = nullptr;
    if (fill) {
        d_filled = mp::pool_alloc(device, s, fill->out_bytes);
        if (!d_filled) {
            mp::pool_free(device, s, d_tab);
            return MP_ERROR_DEVICE_ALLOC;
        }
    }
    if (record_bytes) memcpy((char *)h_tab + tab_bytes, records, record_bytes);
    g_records = record_bytes ? (const char *)d_tab + tab_bytes : nullptr;
    const int out_device = g_out_device >= 0 ? g_out_device : device;
    // Outputs that will live on ANOTHER device (hand-off: this launch writes them over NVLink) are
    // allocated in the order of that device's own work stream -- the stream the receiver will use them
    // on -- and this stream waits for the allocation point through one event.  Allocating a peer pool's
    // memory in the order of a foreign device's stream takes the allocator's slow path on every call
    // (no reuse across devices' streams), which made connected pipelines host-bound.
    cudaStream_t alloc_stream = out_device == device ? s : mp::device_stream(out_device, 1);
    std::vector<void *> fresh(n);
    const bool slab = mp::pool_alloc_many(out_device, alloc_stream, n, out_bytes, fresh.data());
    for (size_t i = 0; i < n; ++i) {
        if (!slab) fresh[i] = mp::pool_alloc(out_device, alloc_stream, out_bytes);
        if (!fresh[i]) {
            for (size_t k = 0; k < i; ++k) mp::pool_free(out_device, alloc_stream, fresh[k]);
            mp::pool_free(device, s, d_tab);
            if (d_filled) mp::pool_free(device, s, d_filled);
            return MP_ERROR_DEVICE_ALLOC;
        }
        h_tab[i] = objs[i]->device_data;
        h_tab[n + i] = fresh[i];
    }
    if (out_device != device) {
        mp::order_after(out_device, alloc_stream, s);
        cudaSetDevice(device);
    }
    MP_CUDA_TRY(cudaMemcpyAsync(d_tab, h_tab, tab_bytes + record_bytes, cudaMemcpyHostToDevice, s));
    if (fill) {
        const char *base = (const char *)d_tab + tab_bytes;
        fill->fill(base, (const unsigned long long *)(base + fill->index_offset), d_filled);
        mp::count_launch();
        g_records = d_filled;
    }
    MPStatus st = launch((const float *const *)d_tab, (float *const *)((void **)d_tab + n), (int)n);
    cudaError_t e = cudaGetLastError();
    if (st == MILLIPYDE_SUCCESS && e != cudaSuccess) {
        mp::record_cuda_error(e, "batched launch", __FILE__, __LINE__);
        st = MP_ERROR_CUDA_RUNTIME;
    }
    for (size_t i = 0; i < n; ++i) {
        if (st == MILLIPYDE_SUCCESS) {
            release_input(device, s, objs[i]);
            objs[i]->device_data = fresh[i];
            objs[i]->nbytes = out_bytes;
            objs[i]->mem_loc = out_device;
        } else {
            mp::pool_free(out_device, alloc_stream, fresh[i]);
        }
    }
    mp::pool_free(device, s, d_tab);
    if (d_filled) mp::pool_free(device, s, d_filled);
    *handled = st == MILLIPYDE_SUCCESS;
    return st;
}

// The per-image programs of a launch, either host-evaluated (`progs`, per_image programs per image) or
// -- when every image shares the template and something in it is drawn -- as a template + the images'
// stream indices for the device-side fill.  Returns true when the device path applies and sets up
// `blob` (what to upload) and `fill`.
bool plan_pw_fill(const std::vector<const Segment *> &segs, const std::vector<uint64_t> &index, int per_image, bool use_pre,
                  uint64_t run_key, cudaStream_t s, std::vector<char> *blob, FillSpec *fill)
{
    if (!g_device_draws.load() || segs.size() < 2 || index.size() != segs.size()) return false;
    const Segment &s0 = *segs[0];
    // a fused Gaussian hands the kernel its folded programs (k_pre, k_post), everything else pre / post
    const bool gauss = s0.kind == Segment::GAUSS_F32;
    auto pre_of = [gauss](const Segment &g) -> const PwTemplate & { return gauss ? g.k_pre_t : g.pre_t; };
    auto post_of = [gauss](const Segment &g) -> const PwTemplate & { return gauss ? g.k_post_t : g.post_t; };
    bool drawn = false;
    for (int i = 0; i < pre_of(s0).n; ++i) drawn = drawn || pre_of(s0).ops[i].stage != kFixedStage;
    for (int i = 0; i < post_of(s0).n; ++i) drawn = drawn || post_of(s0).ops[i].stage != kFixedStage;
    if (!drawn) return false;
    for (size_t i = 0; i < segs.size(); ++i)
        if (!segs[i]->tmpl_ok || memcmp(&pre_of(*segs[i]), &pre_of(s0), sizeof(PwTemplate)) ||
            memcmp(&post_of(*segs[i]), &post_of(s0), sizeof(PwTemplate)))
            return false;
    const size_t n = segs.size();
    const size_t t_bytes = (size_t)per_image * sizeof(PwTemplate);
    blob->resize(t_bytes + n * sizeof(unsigned long long));
    PwTemplate *t = (PwTemplate *)blob->data();
    if (per_image == 2) {
        t[0] = pre_of(s0);
        t[1] = post_of(s0);
    } else {
        t[0] = use_pre ? pre_of(s0) : post_of(s0);
    }
    unsigned long long *idx = (unsigned long long *)(blob->data() + t_bytes);
    for (size_t i = 0; i < n; ++i) idx[i] = index[i];
    fill->index_offset = t_bytes;
    fill->out_bytes = n * per_image * sizeof(PwProgram);
    const int total = (int)n * per_image * kMaxPw;
    fill->fill = [=](const void *d_t, const unsigned long long *d_idx, void *d_out) {
        records_fill_pw_kernel<<<(total + 127) / 128, 128, 0, s>>>((PwProgram *)d_out, (const PwTemplate *)d_t, per_image,
                                                                   d_idx, (int)n, run_key);
    };
    return true;
}

constexpr int kMaxSets = 64;  // images per launch_gauss_stream_sets call (kernels/gaussian_stream.cuh: kGsMaxSets)

// One launch for the Gaussian of a whole same-shape fp32 group (`sigmas` all equal), or one launch
// per 64 images with per-image weight sets when the sigmas were drawn per image (same radius bucket:
// realize() keyed the group on it).
//
// `progs` (optional): 2 n pointwise programs, image i's "before" and "after" the blur (GAUSS_F32
// segments).  They travel as per-image records behind the pointer tables and the kernel applies them
// on the fly -- the chain pointwise -> gaussian -> pointwise is one launch and one HBM round trip.
MPStatus run_gaussian_batch(const std::vector<MPObjData *> &objs, const mp::Img &d, const std::vector<double> &sigmas,
                            int device, cudaStream_t s, bool *handled, const std::vector<PwProgram> *progs = nullptr,
                            const std::vector<const Segment *> *segs = nullptr)
{
    *handled = false;
    const size_t n = objs.size();
    if (n < 2 && !progs) return MILLIPYDE_SUCCESS;
    bool same = true;
    for (size_t i = 1; i < n; ++i) same = same && sigmas[i] == sigmas[0];
    std::vector<GaussParams<float>> gps(same ? 1 : n);
    for (size_t i = 0; i < gps.size(); ++i) {
        if (!(sigmas[i] > 1e-15)) return MILLIPYDE_SUCCESS;
        double w[kGaussMaxRadius + 1];
        const int r = mp::oracle_weights(sigmas[i], w, kGaussMaxRadius);
        const int eff = mp::effective_radius(w, r, mp::kGaussTailEps);
        if (!mp::gauss_stream_supported(d.W, d.C, eff)) return MILLIPYDE_SUCCESS;
        gps[i] = {};
        gps[i].radius = eff;
        for (int k = 0; k <= eff; ++k) gps[i].w[k] = (float)w[k];
    }
    bool same_progs = true;
    if (progs)
        for (size_t i = 1; i < n && same_progs; ++i)
            same_progs = !memcmp(&(*progs)[2 * i], &(*progs)[0], 2 * sizeof(PwProgram));
    const size_t n_records = progs ? (same_progs ? 2 : 2 * n) : 0;
    // programs that differ only by their draws: template + indices up, records filled on the device
    std::vector<char> blob;
    FillSpec fill;
    const bool on_device = progs && !same_progs && segs && g_index &&
                           plan_pw_fill(*segs, *g_index, 2, true, g_run_key, s, &blob, &fill);
    return run_batched(
        objs, objs[0]->nbytes, device, s, handled,
        [&](const float *const *in_tab, float *const *out_tab, int m) {
            if (same && !progs)
                return mp::launch_gauss_stream_batch(device, s, d.H, d.W, d.C, m, nullptr, nullptr, 0, in_tab, out_tab,
                                                     gps[0]);
            const PwProgram *tab = progs ? (const PwProgram *)g_records : nullptr;
            const int pw_stride = (progs && !same_progs) ? 2 : 0;
            const int per_launch = same ? m : kMaxSets;   // one sigma: the weight sets are all alike, no limit
            for (int off = 0; off < m; off += per_launch) {
                const int cnt = m - off < per_launch ? m - off : per_launch;
                MPStatus st = mp::launch_gauss_stream_sets(device, s, d.H, d.W, d.C, cnt, in_tab + off, out_tab + off,
                                                           same ? gps.data() : gps.data() + off, same ? 0 : 1,
                                                           tab ? tab + (size_t)off * pw_stride : nullptr, pw_stride);
                if (st != MILLIPYDE_SUCCESS) return st;
            }
            return (MPStatus)MILLIPYDE_SUCCESS;
        },
        on_device ? (const void *)blob.data() : (progs ? (const void *)progs->data() : nullptr),
        on_device ? blob.size() : n_records * sizeof(PwProgram), on_device ? &fill : nullptr);
}

// The gather segment for one image (tables null) or a same-shape group.
GatherParams gather_params(const Segment &seg, const mp::Img &d)
{
    GatherParams g = {};
    g.out_h = g.rot_h = d.H;
    g.out_w = g.rot_w = g.src_w = d.W;
    g.post = IndexMap{1, 0, 0, 0, seg.flip_post ? -1 : 1, seg.flip_post ? d.W - 1 : 0};
    g.pre = IndexMap{1, 0, 0, 0, seg.flip_pre ? -1 : 1, seg.flip_pre ? d.W - 1 : 0};
    g.has_rotate = seg.has_rotate ? 1 : 0;
    if (seg.has_rotate) g.rp = mp::rotate_params(d.W, d.H, seg.angle);
    g.pw_pre = seg.pre;
    g.pw_post = seg.post;
    return g;
}

// `segs`: one Segment per image (same structure; the angle and the programs may differ) or a single
// one for all.
MPStatus run_gather(const std::vector<MPObjData *> &objs, const std::vector<const Segment *> &segs, const mp::Img &d,
                    int device, cudaStream_t s)
{
    const size_t n = objs.size();
    GatherParams g = gather_params(*segs[0], d);
    std::vector<GatherVar> vars;
    if (segs.size() == n && n >= 2) {
        bool same = true;
        for (size_t i = 1; i < n && same; ++i)
            same = segs[i]->angle == segs[0]->angle && !memcmp(&segs[i]->pre, &segs[0]->pre, sizeof(PwProgram)) &&
                   !memcmp(&segs[i]->post, &segs[0]->post, sizeof(PwProgram));
        if (!same) {
            vars.resize(n);
            for (size_t i = 0; i < n; ++i) {
                const GatherParams gi = gather_params(*segs[i], d);
                vars[i].rp = gi.rp;
                vars[i].pw_pre = gi.pw_pre;
                vars[i].pw_post = gi.pw_post;
            }
        }
    }
    auto launch = [&](const float *const *in_tab, float *const *out_tab, int m) {
        g.in_tab = in_tab;
        g.out_tab = out_tab;
        g.var_tab = vars.empty() ? nullptr : (const GatherVar *)g_records;
        mp::launch_gather_f32(s, d.C, g, m);
        return MILLIPYDE_SUCCESS;
    };
    if (n >= 2) {
        // images that share the segment's template and differ only by their draws (angle, program
        // parameters): one template + the stream indices go up, the records are filled on the device
        std::vector<char> blob;
        FillSpec fill;
        bool on_device = false;
        if (!vars.empty() && g_device_draws.load() && g_index && g_index->size() == n) {
            const Segment &s0 = *segs[0];
            on_device = true;
            for (size_t i = 1; i < n && on_device; ++i)
                on_device = segs[i]->angle_stage == s0.angle_stage && segs[i]->angle_lo == s0.angle_lo &&
                            segs[i]->angle_hi == s0.angle_hi && !memcmp(&segs[i]->pre_t, &s0.pre_t, sizeof(PwTemplate)) &&
                            !memcmp(&segs[i]->post_t, &s0.post_t, sizeof(PwTemplate));
            if (on_device) {
                blob.resize(sizeof(GatherTemplate) + n * sizeof(unsigned long long));
                GatherTemplate *t = (GatherTemplate *)blob.data();
                t->angle_lo = s0.angle_lo;
                t->angle_hi = s0.angle_hi;
                t->angle_stage = s0.angle_stage;
                t->width = d.W;
                t->height = d.H;
                t->pre = s0.pre_t;
                t->post = s0.post_t;
                unsigned long long *idx = (unsigned long long *)(blob.data() + sizeof(GatherTemplate));
                for (size_t i = 0; i < n; ++i) idx[i] = (*g_index)[i];
                fill.index_offset = sizeof(GatherTemplate);
                fill.out_bytes = n * sizeof(GatherVar);
                const int total = (int)n * (1 + 2 * kMaxPw);
                const uint64_t key = g_run_key;
                fill.fill = [=](const void *d_t, const unsigned long long *d_idx, void *d_out) {
                    records_fill_gather_kernel<<<(total + 127) / 128, 128, 0, s>>>((GatherVar *)d_out, (const GatherTemplate *)d_t,
                                                                                   d_idx, (int)n, key);
                };
            }
        }
        bool handled = false;
        MPStatus st = run_batched(objs, objs[0]->nbytes, device, s, &handled, launch,
                                  on_device ? (const void *)blob.data() : (const void *)vars.data(),
                                  on_device ? blob.size() : vars.size() * sizeof(GatherVar), on_device ? &fill : nullptr);
        if (st != MILLIPYDE_SUCCESS || handled) return st;
    }
    for (size_t i = 0; i < n; ++i) {  // one image, or the arena was exhausted
        MPObjData *o = objs[i];
        g = gather_params(*segs[segs.size() == n ? i : 0], d);
        void *fresh = mp::pool_alloc(device, s, o->nbytes);
        if (!fresh) return MP_ERROR_DEVICE_ALLOC;
        g.in = (const float *)o->device_data;
        g.out = (float *)fresh;
        mp::launch_gather_f32(s, d.C, g, 1);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            mp::record_cuda_error(e, "gather launch", __FILE__, __LINE__);
            mp::pool_free(device, s, fresh);
            return MP_ERROR_CUDA_RUNTIME;
        }
        release_input(device, s, o);
        o->device_data = fresh;
    }
    return MILLIPYDE_SUCCESS;
}

// One segment for a group of images that share the layout and the segment's signature; `segs[i]`
// is image i's own segment.  Parameters that differ between the images (random_* draws) travel as
// per-image records, so the group is one launch.
void run_segment_group(mp_pipeline *p, const std::vector<MPObjData *> &objs, const std::vector<const Segment *> &segs,
                       int device, cudaStream_t s)
{
    const size_t n = objs.size();
    if (!n) return;
    const Segment &seg = *segs[0];
    mp::Img cur;
    const bool f32 = mp::describe(objs[0], &cur) && cur.fam == mp::FAM_F32;
    if (n >= 2) cudaSetDevice(device);

    if (seg.kind == Segment::SINGLE && seg.single->kind == OP_GAUSSIAN && g_fusion.load() && f32) {
        std::vector<double> sigmas(n);
        for (size_t i = 0; i < n; ++i) sigmas[i] = segs[i]->single->a[0];
        bool handled = false;
        note_status(p, run_gaussian_batch(objs, cur, sigmas, device, s, &handled));
        if (handled) return;
    }
    if (seg.kind == Segment::GAUSS_F32 && f32) {
        std::vector<double> sigmas(n);
        std::vector<PwProgram> progs(2 * n);
        for (size_t i = 0; i < n; ++i) {
            sigmas[i] = segs[i]->single->a[0];
            progs[2 * i] = segs[i]->k_pre;
            progs[2 * i + 1] = segs[i]->k_post;
        }
        bool handled = false;
        note_status(p, run_gaussian_batch(objs, cur, sigmas, device, s, &handled, &progs, &segs));
        if (handled) return;
    }
    if ((seg.kind == Segment::PW_F32 || seg.kind == Segment::GREY_F32) && n >= 2 && f32 &&
        (seg.kind == Segment::PW_F32 || (cur.C >= 3 && objs[0]->ndims == 3))) {
        // one launch for the whole group (pointer tables); the grey variant also rewrites headers
        const bool grey = seg.kind == Segment::GREY_F32;
        const size_t out_bytes = grey ? cur.npix * 4 : objs[0]->nbytes;
        std::vector<PwProgram> progs;  // per-image records: [i] (pw) or [2i], [2i+1] (grey)
        bool same = true;
        for (size_t i = 1; i < n && same; ++i)
            same = !memcmp(&segs[i]->pre, &seg.pre, sizeof(PwProgram)) && !memcmp(&segs[i]->post, &seg.post, sizeof(PwProgram));
        if (!same)
            for (size_t i = 0; i < n; ++i) {
                progs.push_back(segs[i]->pre);
                if (grey) progs.push_back(segs[i]->post);
            }
        bool handled = false;
        std::vector<char> blob;
        FillSpec fill;
        const bool on_device = !same && g_index && plan_pw_fill(segs, *g_index, grey ? 2 : 1, true, g_run_key, s, &blob, &fill);
        MPStatus st = run_batched(
            objs, out_bytes, device, s, &handled,
            [&](const float *const *in_tab, float *const *out_tab, int m) {
                const PwProgram *tab = progs.empty() ? nullptr : (const PwProgram *)g_records;
                if (grey) mp::launch_grey_f32_batch(s, cur, seg.pre, seg.post, in_tab, out_tab, m, tab);
                else mp::launch_pw_f32_batch(s, cur, seg.pre, in_tab, out_tab, m, tab);
                return MILLIPYDE_SUCCESS;
            },
            on_device ? (const void *)blob.data() : (const void *)progs.data(),
            on_device ? blob.size() : progs.size() * sizeof(PwProgram), on_device ? &fill : nullptr);
        note_status(p, st);
        if (handled) {
            if (grey)
                for (MPObjData *o : objs) {  // header rewrite of mpimg_color_to_greyscale
                    o->ndims = 2;
                    o->type = MP_NPY_FLOAT;
                    o->dims[2] = cur.W * 4;
                    o->dims[3] = 4;
                }
            return;
        }
        if (st != MILLIPYDE_SUCCESS) return;
    }
    if (seg.kind == Segment::SINGLE && seg.single->kind == OP_FLIPLR && n >= 2 && g_fusion.load() &&
        mp::describe(objs[0], &cur) && mp::fliplr_batch_supported(cur)) {
        bool handled = false;
        MPStatus st = run_batched(objs, objs[0]->nbytes, device, s, &handled,
                                  [&](const float *const *in_tab, float *const *out_tab, int m) {
                                      mp::launch_fliplr_batch(s, cur, (const void *const *)in_tab,
                                                              (void *const *)out_tab, m);
                                      return MILLIPYDE_SUCCESS;
                                  });
        note_status(p, st);
        if (handled || st != MILLIPYDE_SUCCESS) return;
    }
    if (seg.kind == Segment::SINGLE && seg.single->kind == OP_TRANSPOSE && n >= 2 && g_fusion.load() &&
        mp::describe(objs[0], &cur) && mp::transpose_batch_supported(cur)) {
        bool handled = false;
        MPStatus st = run_batched(objs, objs[0]->nbytes, device, s, &handled,
                                  [&](const float *const *in_tab, float *const *out_tab, int m) {
                                      mp::launch_transpose_batch(s, cur, (const void *const *)in_tab,
                                                                 (void *const *)out_tab, m);
                                      return MILLIPYDE_SUCCESS;
                                  });
        note_status(p, st);
        if (handled) {
            const int pix_bytes = cur.C * cur.esize;
            for (MPObjData *o : objs) {  // header rewrite of mpimg_transpose
                o->dims[0] = cur.W;
                o->dims[1] = cur.H;
                o->dims[o->ndims] = cur.H * pix_bytes;
                o->dims[o->ndims + 1] = pix_bytes;
            }
        }
        if (handled || st != MILLIPYDE_SUCCESS) return;
    }
    if (seg.kind == Segment::GATHER_F32 && f32) {
        cudaSetDevice(device);
        note_status(p, run_gather(objs, segs, cur, device, s));
        return;
    }
    for (size_t i = 0; i < n; ++i) {
        if (segs[i]->kind != Segment::GATHER_F32) {  // eager ops free their input: own it first
            MPStatus st = materialize(device, s, objs[i]);
            note_status(p, st);
            if (st != MILLIPYDE_SUCCESS) continue;
        }
        note_status(p, run_segment_on(objs[i], *segs[i]));
    }
}

// The layout part of a grouping key: images must agree on it to share a launch.
std::string layout_key(const MPObjData *o)
{
    char hdr[64];
    const int c = o->ndims == 3 ? o->dims[2] : 1;
    snprintf(hdr, sizeof hdr, "%d:%d:%d:%d:%d|", o->type, o->ndims, o->ndims > 0 ? o->dims[0] : 0,
             o->ndims > 1 ? o->dims[1] : 0, c);
    return hdr;
}

constexpr size_t kHandoffChunk = 32;  // images per hand-off chunk of a cross-device pipeline pair

struct ShardTask {
    mp_pipeline *pipe;
    std::vector<MPObjData *> objs;
    std::vector<uint64_t> index;  // stream index of every object (position in the submitted array + the pipeline's base)
    int device;
    bool views;  // objs borrow their buffers (mppipe_run_views)
    cudaEvent_t after = nullptr;  // work of the sending shard that wrote into this device's memory
};

void submit_shard(mp_pipeline *p, std::vector<MPObjData *> objs, std::vector<uint64_t> index, int device, bool views = false,
                  cudaEvent_t after = nullptr);

void shard_worker(void *arg)
{
    ShardTask *t = (ShardTask *)arg;
    mp_pipeline *p = t->pipe;
    const int device = t->device;
    const size_t n = t->objs.size();
    cudaSetDevice(device);
    cudaStream_t batch_stream = mp::device_stream(device, 1);
    // MILLIPYDE_TRACE=1: one stderr line per shard with the host time spent enqueueing it and the time
    // spent waiting for the device afterwards (tells host-bound batches from device-bound ones)
    static const bool trace = [] { const char *e = getenv("MILLIPYDE_TRACE"); return e && *e && *e != '0'; }();
    const auto t_start = std::chrono::steady_clock::now();

    DeviceArenas &da = g_arenas[device];
    std::unique_lock<std::mutex> enqueue_lock(da.enqueue);
    Arena &arena = da.ring[da.next++ % kArenaRing];
    const unsigned long long launches0 = mpdev_launch_count();
    if (arena.in_use) cudaEventSynchronize(arena.done);   // its previous shard's copies have been read
    const size_t want = n * (sizeof(void *) * 2 + sizeof(GatherVar) + 64) * (p->stages.size() + 1) + 4096;
    if (arena.cap < want) {
        if (arena.base) cudaFreeHost(arena.base);
        arena.base = nullptr;
        arena.cap = 0;
        if (cudaHostAlloc((void **)&arena.base, want * 2, cudaHostAllocPortable) == cudaSuccess) arena.cap = want * 2;
        else (void)cudaGetLastError();
    }
    if (!arena.done && cudaEventCreateWithFlags(&arena.done, cudaEventDisableTiming) != cudaSuccess) {
        (void)cudaGetLastError();
        arena.done = nullptr;
    }
    arena.used = 0;
    g_arena = &arena;

    if (t->after) {  // the sender's kernels wrote some of these images straight into this device's memory
        cudaStreamWaitEvent(batch_stream, t->after, 0);
        cudaEventDestroy(t->after);
    }
    std::unordered_set<MPObjData *> borrowed;
    if (t->views) borrowed.insert(t->objs.begin(), t->objs.end());
    g_borrowed = t->views ? &borrowed : nullptr;

    // 1. bring every object onto this device and onto the shard's stream.  The stream hand-over is
    // ONE event per distinct source stream, not one per object (a Generator batch of 2,048 views all
    // come from stream 0: 2,048 event create / record / wait / destroy round trips through the driver
    // were most of the host time of a batch)
    // Two passes: every cross-device move first (a move makes the object's new stream wait for its
    // copy), THEN the hand-over events -- an event recorded on a stream between two moves would order
    // the shard's stream after the first copy only, and the launches below would race the others
    // (seen as wrong images in a cycling run over 4 devices).
    {
        for (size_t i = 0; i < n; ++i) {
            MPObjData *o = t->objs[i];
            o->pinned = MP_TRUE;
            if (o->mem_loc != device && o->device_data) mpobj_change_device(o, device);
        }
        cudaStream_t seen[8];
        int n_seen = 0;
        for (size_t i = 0; i < n; ++i) {
            MPObjData *o = t->objs[i];
            cudaStream_t old = mp::stream_of(o);
            if (o->device_data && old && old != batch_stream) {
                bool known = false;
                for (int k = 0; k < n_seen; ++k) known = known || seen[k] == old;
                if (!known) {
                    mp::order_after(o->mem_loc, old, batch_stream);
                    if (n_seen < 8) seen[n_seen++] = old;
                }
            }
            o->stream = (void *)batch_stream;
        }
    }

    // A chain that hands its images to a pipeline on ANOTHER device writes the result of each
    // image's last segment straight into that device's memory (batched launches only; whatever
    // stays here is moved by the receiver with cudaMemcpyPeerAsync as before).
    const int remote = (p->receiver && p->receiver->device != device && mpdev_is_valid_device(p->receiver->device) &&
                        g_fusion.load() && (mpdev_can_use_peer(device, p->receiver->device) != 0))
                           ? p->receiver->device
                           : -1;
    // With a receiver on another device the shard goes through in chunks: while this device works
    // on chunk k + 1 the receiver already has chunk k (its launches are still whole-chunk launches).
    // A large shard without a receiver also goes through in chunks: the host work of chunk k + 1 --
    // realise, compile, group, allocate, tables -- overlaps the kernels of chunk k instead of keeping
    // the device idle until the whole shard is prepared (1,024 images: ~3 ms).  The first chunk is
    // kShardChunk images, every further one twice the one before: the device starts early, and a long
    // look-ahead still runs as a few large launches (chunks of 256 throughout made a 2,048-image
    // Generator batch five times the launches and cost it 6-13 %; one chunk costs config 3 15 %).
    // (MILLIPYDE_SHARD_CHUNK overrides the first chunk's size: A/B measurements)
    static const size_t kShardChunk = [] {
        const char *e = getenv("MILLIPYDE_SHARD_CHUNK");
        const long v = e ? atol(e) : 0;
        return (size_t)(v > 0 ? v : 256);
    }();
    const bool handoff_chunks = remote >= 0 && n > 2 * kHandoffChunk;
    size_t chunk = handoff_chunks ? kHandoffChunk : (n >= 3 * kShardChunk ? kShardChunk : (n ? n : 1));
    for (size_t c0 = 0, c1 = 0; c0 < n; c0 = c1, chunk = handoff_chunks ? chunk : 2 * chunk) {
        c1 = c0 + chunk < n ? c0 + chunk : n;
        if (!handoff_chunks && n - c1 < kShardChunk / 2) c1 = n;   // no sliver at the end
        // 2. coin flips / random draws per image, then the fusion pass on what survived
        std::vector<std::vector<Stage>> realized(n);
        std::vector<std::vector<Segment>> segs(n);
        size_t rounds = 0;
        for (size_t i = c0; i < c1; ++i) {
            realize(p, t->index[i], &realized[i]);
            std::vector<const Stage *> ops;
            for (const Stage &st : realized[i]) ops.push_back(&st);
            mp::Img d;
            const bool known = mp::describe(t->objs[i], &d);
            segs[i] = compile(ops, known ? d.fam : mp::FAM_U8_OTHER, known ? d.C : 0);
            if (segs[i].size() > rounds) rounds = segs[i].size();
        }

        // 3. round r runs every image's r-th segment; the images of a round are regrouped by (current
        // layout, segment signature), so images whose coins fell differently part and meet again, and
        // a chain of random_* stages still costs one launch per (round, kernel), not per image.
        for (size_t r = 0; r < rounds; ++r) {
            std::map<std::string, std::vector<size_t>> groups;
            for (size_t i = c0; i < c1; ++i)
                if (r < segs[i].size()) {
                    const bool last = remote >= 0 && r + 1 == segs[i].size();
                    groups[layout_key(t->objs[i]) + signature(segs[i][r]) + (last ? "|L" : "")].push_back(i);
                }
            for (auto &g : groups) {
                std::vector<MPObjData *> objs;
                std::vector<const Segment *> sp;
                std::vector<uint64_t> idx;
                for (size_t i : g.second) {
                    objs.push_back(t->objs[i]);
                    sp.push_back(&segs[i][r]);
                    idx.push_back(t->index[i]);
                }
                const bool last = remote >= 0 && r + 1 == segs[g.second[0]].size();
                g_out_device = last ? remote : -1;
                g_index = &idx;
                g_run_key = p->run_key;
                run_segment_group(p, objs, sp, device, batch_stream);
                g_index = nullptr;
                g_out_device = -1;
            }
            p->segments.fetch_add((int)groups.size());
        }

        // 4b. a view no segment rewrote still borrows: deep-copy it
        if (t->views)
            for (size_t i = c0; i < c1; ++i) note_status(p, materialize(device, batch_stream, t->objs[i]));

        // 5. hand the results on, or finish
        if (p->receiver) {
            // Images still here: the receiver's worker orders itself after this stream through the
            // objects' events and moves them.  Images already written into the receiver's memory: one
            // event for the whole shard, and they are handed over as belonging to its stream.
            cudaEvent_t after = nullptr;
            if (remote >= 0) {
                bool any = false;
                for (size_t i = c0; i < c1; ++i) any = any || (t->objs[i]->mem_loc == remote && t->objs[i]->device_data);
                if (any) {
                    cudaSetDevice(device);
                    if (cudaEventCreateWithFlags(&after, cudaEventDisableTiming) == cudaSuccess &&
                        cudaEventRecord(after, batch_stream) == cudaSuccess) {
                        for (size_t i = c0; i < c1; ++i)
                            if (t->objs[i]->mem_loc == remote) t->objs[i]->stream = (void *)mp::device_stream(remote, 1);
                    } else {
                        (void)cudaGetLastError();
                        if (after) cudaEventDestroy(after);
                        after = nullptr;
                        cudaStreamSynchronize(batch_stream);  // no event: hand over finished work
                        for (size_t i = c0; i < c1; ++i)
                            if (t->objs[i]->mem_loc == remote) t->objs[i]->stream = (void *)mp::device_stream(remote, 1);
                    }
                }
            }
            submit_shard(p->receiver, std::vector<MPObjData *>(t->objs.begin() + c0, t->objs.begin() + c1),
                         std::vector<uint64_t>(t->index.begin() + c0, t->index.begin() + c1), p->receiver->device, false,
                         after);
        }
    }
    if (t->views) g_borrowed = nullptr;

    // Everything is enqueued: mark the end, let the next shard of this device start enqueueing, and
    // wait for OUR work only (the stream may already hold the next shard's).
    g_arena = nullptr;
    const unsigned long long launched = mpdev_launch_count() - launches0;
    cudaEvent_t done = arena.done;
    arena.in_use = done && cudaEventRecord(done, batch_stream) == cudaSuccess;
    const bool have_done = arena.in_use;
    enqueue_lock.unlock();
    const auto t_enqueued = std::chrono::steady_clock::now();
    cudaError_t e = have_done ? cudaEventSynchronize(done) : cudaStreamSynchronize(batch_stream);
    if (trace) {
        const auto t_done = std::chrono::steady_clock::now();
        fprintf(stderr, "[millipyde] shard dev=%d images=%zu enqueue=%.3f ms wait=%.3f ms\n", device, n,
                std::chrono::duration<double, std::milli>(t_enqueued - t_start).count(),
                std::chrono::duration<double, std::milli>(t_done - t_enqueued).count());
    }
    if (e != cudaSuccess) {
        mp::record_cuda_error(e, "cudaStreamSynchronize(shard)", __FILE__, __LINE__);
        note_status(p, MP_ERROR_CUDA_RUNTIME);
    }
    if (!p->receiver) {
        for (size_t i = 0; i < n; ++i) {
            t->objs[i]->pinned = MP_FALSE;
            t->objs[i]->stream = (void *)mp::device_stream(device, 0);  // idle: no event needed
        }
    }
    p->launches.fetch_add(launched);
    delete t;
    {
        std::lock_guard<std::mutex> lk(p->mux);
        --p->pending;
    }
    p->cv.notify_all();
}

void submit_shard(mp_pipeline *p, std::vector<MPObjData *> objs, std::vector<uint64_t> index, int device, bool views,
                  cudaEvent_t after)
{
    if (objs.empty()) return;
    {
        std::lock_guard<std::mutex> lk(p->mux);
        p->in_flight.push_back(device);
        ++p->pending;
    }
    ShardTask *t = new ShardTask{p, std::move(objs), std::move(index), device, views, after};
    mpdev_submit_work(device, shard_worker, t);
}

void reset_counters(mp_pipeline *p)
{
    for (mp_pipeline *q = p; q; q = q->receiver) {
        q->launches.store(0);
        q->segments.store(0);
        q->status.store(MILLIPYDE_SUCCESS);
        std::lock_guard<std::mutex> lk(q->mux);
        q->in_flight.clear();
    }
}

}  // namespace

extern "C" {

MPPipeline *mppipe_create(const MPRunnable *stages, int num_stages, int device_id)
{
    if (num_stages < 0 || (num_stages > 0 && !stages)) return nullptr;
    mp_pipeline *p = new (std::nothrow) mp_pipeline();
    if (!p) return nullptr;
    p->device = device_id;
    for (int i = 0; i < num_stages; ++i) {
        Stage s = {};
        size_t bytes = 0;
        s.kind = classify(stages[i].func, &bytes);
        s.func = stages[i].func;
        s.probability = stages[i].probability;
        s.args = stages[i].args;
        if (bytes && !stages[i].args) {
            s.kind = OP_NULL;  // an operator that needs arguments but got none: skipped like func == NULL
        } else if (bytes) {
            s.args = malloc(bytes);
            memcpy(s.args, stages[i].args, bytes);
            const double *a = (const double *)s.args;
            for (size_t k = 0; k < bytes / sizeof(double) && k < 5; ++k) s.a[k] = a[k];
        }
        p->stages.push_back(s);
    }
    return p;
}

void mppipe_destroy(MPPipeline *p)
{
    if (!p) return;
    for (Stage &s : p->stages) {
        size_t bytes = 0;
        if (s.kind != OP_FOREIGN && s.kind != OP_NULL) {
            classify(s.func, &bytes);
            if (bytes) free(s.args);
        }
    }
    delete p;
}

int mppipe_get_device(const MPPipeline *p) { return p ? p->device : DEVICE_LOC_NO_AFFINITY; }
void mppipe_set_device(MPPipeline *p, int device_id)
{
    if (p) p->device = device_id;
}

// Device assignment rules of PyGPUPipeline_connect_to (src/gpupipeline.c:186-218).
void mppipe_connect(MPPipeline *self, MPPipeline *receiver)
{
    if (!self || !receiver) return;
    const int recommended = mpdev_get_recommended_device();
    const int alternative = mpdev_get_alternative_device(recommended);
    if (alternative == DEVICE_LOC_NO_AFFINITY) {
        self->device = recommended;
        receiver->device = recommended;
    } else if (self->device == DEVICE_LOC_NO_AFFINITY && receiver->device == DEVICE_LOC_NO_AFFINITY) {
        self->device = recommended;
        receiver->device = alternative;
    } else if (self->device == DEVICE_LOC_NO_AFFINITY) {
        self->device = mpdev_get_alternative_device(receiver->device);
    } else if (receiver->device == DEVICE_LOC_NO_AFFINITY) {
        receiver->device = mpdev_get_alternative_device(self->device);
    }
    self->receiver = receiver;
}

static MPStatus submit_impl(MPPipeline *p, MPObjData **objs, int n, bool views);

MPStatus mppipe_submit(MPPipeline *p, MPObjData **objs, int n) { return submit_impl(p, objs, n, false); }

MPStatus mppipe_submit_views(MPPipeline *p, MPObjData **views, int n)
{
    MPStatus st = submit_impl(p, views, n, true);
    if (st != MILLIPYDE_SUCCESS) {  // nothing ran: the views must not keep (and later free) what they borrow
        for (int i = 0; views && i < n; ++i)
            if (views[i]) views[i]->device_data = NULL;
    }
    return st;
}

MPStatus mppipe_run_views(MPPipeline *p, MPObjData **views, int n)
{
    MPStatus st = mppipe_submit_views(p, views, n);
    if (st != MILLIPYDE_SUCCESS) return st;
    return mppipe_wait(p);
}

static MPStatus submit_impl(MPPipeline *p, MPObjData **objs, int n, bool views)
{
    if (!p || n < 0 || (n > 0 && !objs)) return MP_ERROR_INVALID_ARGUMENT;
    MPStatus st = mp::ensure_initialized();
    if (st != MILLIPYDE_SUCCESS) return st;
    reset_counters(p);

    int device = p->device;
    bool cycle = false;
    if (device == DEVICE_LOC_NO_AFFINITY) {
        const int target = mpdev_get_target_device();
        if (target != DEVICE_LOC_NO_AFFINITY) {
            device = target;
        } else {
            cycle = true;
            device = mpdev_get_recommended_device();
        }
    }
    if (views) {
        // a view stays where the buffer it borrows lives: one shard per device that holds views (a
        // Generator spreading its stream borrows from per-device replicas of its inputs)
        cycle = false;
        for (int i = 0; i < n; ++i)
            if (!objs[i] || !objs[i]->device_data || !mpdev_is_valid_device(objs[i]->mem_loc)) return MP_ERROR_INVALID_ARGUMENT;
        if (n > 0) device = objs[0]->mem_loc;
    }
    if (!mpdev_is_valid_device(device)) return GPUPIPELINE_ERROR_UNUSABLE_DEVICE;
    p->cycled = cycle;
    p->soft_wait = true;   // every shard (and every receiver shard) is counted in `pending`: wait on that
    for (mp_pipeline *q = p->receiver; q; q = q->receiver)
        if (!mpdev_is_valid_device(q->device)) q->device = device;

    // Image i -> device: blocks of THREADS_PER_DEVICE images, round-robin over the valid
    // devices starting at the recommended one (src/gpupipeline.c:267-283).
    // one random-source key per run, shared down the receiver chain (stages are numbered per pipeline,
    // so a receiver offsets its stage numbers through its own key)
    {
        const uint64_t key = p->key_held ? p->run_key : mp::next_run_key();
        uint64_t salt = 0;
        for (mp_pipeline *q = p; q; q = q->receiver) {
            q->run_key = key + 0x9E3779B97F4A7C15ull * salt++;
            q->index_base = p->index_base;
        }
    }
    std::map<int, std::vector<MPObjData *>> shards;
    std::map<int, std::vector<uint64_t>> shard_index;
    int cur = device;
    for (int i = 0; i < n; ++i) {
        if (objs[i]) {
            const int dev = views ? objs[i]->mem_loc : cur;
            shards[dev].push_back(objs[i]);
            shard_index[dev].push_back(p->index_base + (uint64_t)i);
        }
        if (cycle && ((i + 1) % THREADS_PER_DEVICE == 0)) cur = mpdev_get_next_device(cur);
    }
    for (auto &kv : shards) submit_shard(p, std::move(kv.second), std::move(shard_index[kv.first]), kv.first, views);
    return MILLIPYDE_SUCCESS;
}

MPStatus mppipe_wait(MPPipeline *p)
{
    if (!p) return MP_ERROR_INVALID_ARGUMENT;
    if (p->soft_wait) {
        // every shard's worker returns only after its device work is complete, and a sender registers
        // its receiver's shards before it finishes: own shards first, then down the chain
        for (mp_pipeline *q = p; q; q = q->receiver) {
            std::unique_lock<std::mutex> lk(q->mux);
            q->cv.wait(lk, [q] { return q->pending == 0; });
        }
    } else {
        // own device first, then down the receiver chain: by the time a device's pool is
        // drained its workers have enqueued their hand-offs (src/gpupipeline.c:293-306)
        for (mp_pipeline *q = p; q; q = q->receiver) {
            std::vector<int> devs;
            {
                std::lock_guard<std::mutex> lk(q->mux);
                devs = q->in_flight;
            }
            for (int d : devs) mpdev_hard_synchronize(d);
        }
    }
    MPStatus st = MILLIPYDE_SUCCESS;
    for (mp_pipeline *q = p; q; q = q->receiver) {
        if (q->status.load() != MILLIPYDE_SUCCESS && st == MILLIPYDE_SUCCESS) st = (MPStatus)q->status.load();
        if (q != p) {
            p->launches.fetch_add(q->launches.load());
            p->segments.fetch_add(q->segments.load());
        }
    }
    return st;
}

MPStatus mppipe_run(MPPipeline *p, MPObjData **objs, int n)
{
    MPStatus st = mppipe_submit(p, objs, n);
    if (st != MILLIPYDE_SUCCESS) return st;
    return mppipe_wait(p);
}

int mppipe_plan(const MPPipeline *p, int typenum, int channels, char *buf, int cap)
{
    if (!p || !buf || cap < 1 || channels < 1) return -1;
    mp::Family fam;
    switch (typenum) {
        case MP_NPY_UBYTE: fam = channels == 4 ? mp::FAM_RGBA8 : mp::FAM_U8_OTHER; break;
        case MP_NPY_DOUBLE: fam = channels == 1 ? mp::FAM_F64 : mp::FAM_F64_OTHER; break;
        case MP_NPY_FLOAT:
            if (channels != 1 && channels != 3 && channels != 4) return -1;
            fam = mp::FAM_F32;
            break;
        default: return -1;
    }
    mp_pipeline all = {};   // the same stages with every coin forced to "run"
    all.stages = p->stages;
    for (Stage &st : all.stages) st.probability = -1;
    std::vector<Stage> realized;
    all.run_key = mp::peek_run_key();
    realize(&all, 0, &realized);
    std::vector<const Stage *> ops;
    for (const Stage &st : realized) ops.push_back(&st);
    const std::vector<Segment> segs = compile(ops, fam, channels);
    static const char *const kOp[] = {"rgb2grey", "transpose", "gaussian", "fliplr", "rotate", "brightness",
                                      "adjust_gamma", "colorize", "elementwise", "random", "foreign", "null"};
    static const char *const kPw[] = {"none", "brightness", "adjust_gamma", "colorize", "add", "multiply", "power", "clip"};
    auto prog = [](const PwProgram &g) {
        std::string t;
        for (int i = 0; i < g.n; ++i) t += (i ? "," : "") + std::string(kPw[g.ops[i].kind & 7]);
        return t;
    };
    std::string out;
    for (const Segment &g : segs) {
        if (!out.empty()) out += ";";
        switch (g.kind) {
            case Segment::SINGLE: out += kOp[g.single->kind]; break;
            case Segment::PW_F32: out += "pw(" + prog(g.pre) + ")"; break;
            case Segment::GREY_F32: out += "grey(" + prog(g.pre) + "|" + prog(g.post) + ")"; break;
            case Segment::PW_RGBA8: {
                std::string t;
                for (int i = 0; i < g.u8.n; ++i) t += (i ? "," : "") + std::string(kPw[g.u8.ops[i].kind & 3]);
                out += "u8(" + t + ")";
                break;
            }
            case Segment::GAUSS_F32: out += "gauss(" + prog(g.pre) + "|" + prog(g.post) + ")"; break;
            case Segment::GATHER_F32:
                out += std::string("gather(") + (g.flip_pre ? "fliplr" : "-") + "," + (g.has_rotate ? "rotate" : "-") + "," +
                       (g.flip_post ? "fliplr" : "-") + ";" + prog(g.pre) + "|" + prog(g.post) + ")";
                break;
        }
    }
    if ((int)out.size() + 1 > cap) return -1;
    memcpy(buf, out.c_str(), out.size() + 1);
    return (int)segs.size();
}

void mppipe_set_index_base(MPPipeline *p, unsigned long long first_index)
{
    if (p) p->index_base = first_index;
}
unsigned long long mppipe_last_run_key(const MPPipeline *p) { return p ? p->run_key : 0; }
void mppipe_hold_run_key(MPPipeline *p)
{
    if (!p || p->key_held) return;
    p->run_key = mp::next_run_key();
    p->key_held = true;
}
void mppipe_set_device_draws(int enabled) { g_device_draws.store(enabled ? 1 : 0); }
int mppipe_get_device_draws(void) { return g_device_draws.load(); }

void mppipe_set_fusion(int enabled) { g_fusion.store(enabled ? 1 : 0); }
int mppipe_get_fusion(void) { return g_fusion.load(); }
unsigned long long mppipe_last_launches(const MPPipeline *p) { return p ? p->launches.load() : 0; }
int mppipe_last_segments(const MPPipeline *p) { return p ? p->segments.load() : 0; }

/* ------------------------------------------------------- host streaming path */

}  // extern "C"

namespace {

struct HostTask {
    mp_pipeline *pipe;
    int device;
    const void *const *host_in;
    void *const *host_out;
    size_t out_capacity;
    MPHostResult *results;
    std::vector<int> indices;
    int ndims;
    long shape[3];
    int typenum;
    size_t in_bytes;
};

void host_worker(void *arg)
{
    HostTask *t = (HostTask *)arg;
    mp_pipeline *p = t->pipe;
    const int device = t->device;
    cudaSetDevice(device);
    const unsigned long long launches0 = mpdev_launch_count();
    const int depth = THREADS_PER_DEVICE;  // one image in flight per work stream

    struct Slot {
        MPObjData obj;
        int dims[6];
        cudaEvent_t done;
        bool busy;
    } slots[THREADS_PER_DEVICE];
    for (int k = 0; k < depth; ++k) {
        memset(&slots[k].obj, 0, sizeof(MPObjData));
        slots[k].busy = false;
        cudaEventCreateWithFlags(&slots[k].done, cudaEventDisableTiming);
    }

    for (size_t n = 0; n < t->indices.size(); ++n) {
        const int i = t->indices[n];
        Slot &sl = slots[n % depth];
        cudaStream_t s = mp::device_stream(device, 1 + (int)(n % depth));
        if (sl.busy) {
            cudaEventSynchronize(sl.done);
            sl.busy = false;
        }
        MPObjData *o = &sl.obj;
        o->ndims = t->ndims;
        o->dims = sl.dims;
        size_t stride = t->in_bytes;
        for (int k = 0; k < t->ndims; ++k) {
            stride /= (size_t)t->shape[k];
            sl.dims[k] = (int)t->shape[k];
            sl.dims[t->ndims + k] = (int)stride;
        }
        o->type = t->typenum;
        o->mem_loc = device;
        o->stream = (void *)s;
        o->pinned = MP_TRUE;
        o->nbytes = t->in_bytes;
        o->device_data = mp::pool_alloc(device, s, t->in_bytes);
        MPStatus st = o->device_data ? MILLIPYDE_SUCCESS : MP_ERROR_DEVICE_ALLOC;
        if (st == MILLIPYDE_SUCCESS) st = mpobj_upload_async(o, t->host_in[i], t->in_bytes);

        std::vector<Stage> mine;
        realize(p, p->index_base + (uint64_t)i, &mine);
        std::vector<const Stage *> ops;
        for (const Stage &sg : mine) ops.push_back(&sg);
        if (st == MILLIPYDE_SUCCESS) {
            mp::Img d;
            const bool known = mp::describe(o, &d);
            std::vector<Segment> segs = compile(ops, known ? d.fam : mp::FAM_U8_OTHER, known ? d.C : 0);
            p->segments.fetch_add((int)segs.size());
            for (const Segment &seg : segs) {
                st = run_segment_on(o, seg);
                if (st != MILLIPYDE_SUCCESS) break;
            }
        }
        MPHostResult &r = t->results[i];
        r.status = (int)st;
        r.ndims = o->ndims;
        for (int k = 0; k < 3; ++k) r.shape[k] = k < o->ndims ? o->dims[k] : 0;
        r.type = o->type;
        r.nbytes = o->nbytes;
        if (st == MILLIPYDE_SUCCESS) {
            if (o->nbytes > t->out_capacity) {
                r.status = (int)MP_ERROR_INVALID_ARGUMENT;
            } else {
                st = mpobj_download_async(o, t->host_out[i], o->nbytes);
                if (st != MILLIPYDE_SUCCESS) r.status = (int)st;
            }
        }
        note_status(p, (MPStatus)r.status);
        if (o->device_data) mp::pool_free(device, s, o->device_data);
        o->device_data = NULL;
        cudaEventRecord(sl.done, s);
        sl.busy = true;
    }
    for (int k = 0; k < depth; ++k) {
        if (slots[k].busy) cudaEventSynchronize(slots[k].done);
        cudaEventDestroy(slots[k].done);
    }
    p->launches.fetch_add(mpdev_launch_count() - launches0);
    delete t;
}

}  // namespace

extern "C" MPStatus mppipe_run_host(MPPipeline *p, const void *const *host_in, void *const *host_out,
                                    size_t out_capacity, MPHostResult *results, int n, int ndims,
                                    const long *shape, int typenum)
{
    if (!p || n < 0 || ndims < 2 || ndims > 3 || !shape || (n > 0 && (!host_in || !host_out || !results)))
        return MP_ERROR_INVALID_ARGUMENT;
    if (p->receiver) return MP_ERROR_INVALID_ARGUMENT;   // the host stream has no hand-off: say so, do not drop it
    MPStatus st = mp::ensure_initialized();
    if (st != MILLIPYDE_SUCCESS) return st;
    reset_counters(p);
    if (!p->key_held) p->run_key = mp::next_run_key();
    size_t es;
    switch (typenum) {
        case MP_NPY_UBYTE: es = 1; break;
        case MP_NPY_FLOAT: es = 4; break;
        case MP_NPY_DOUBLE: es = 8; break;
        default: return MP_ERROR_UNSUPPORTED_LAYOUT;
    }
    size_t in_bytes = es;
    for (int k = 0; k < ndims; ++k) in_bytes *= (size_t)shape[k];

    int device = p->device;
    bool cycle = false;
    if (device == DEVICE_LOC_NO_AFFINITY) {
        const int target = mpdev_get_target_device();
        if (target != DEVICE_LOC_NO_AFFINITY) device = target;
        else {
            cycle = true;
            device = mpdev_get_recommended_device();
        }
    }
    if (!mpdev_is_valid_device(device)) return GPUPIPELINE_ERROR_UNUSABLE_DEVICE;
    p->cycled = cycle;
    p->soft_wait = false;   // host_worker tasks are not counted in `pending`

    std::map<int, std::vector<int>> shards;
    int cur = device;
    for (int i = 0; i < n; ++i) {
        shards[cur].push_back(i);
        if (cycle && ((i + 1) % THREADS_PER_DEVICE == 0)) cur = mpdev_get_next_device(cur);
    }
    for (auto &kv : shards) {
        HostTask *t = new HostTask{p, kv.first, host_in, host_out, out_capacity, results, kv.second, ndims,
                                   {0, 0, 0}, typenum, in_bytes};
        for (int k = 0; k < ndims; ++k) t->shape[k] = shape[k];
        {
            std::lock_guard<std::mutex> lk(p->mux);
            p->in_flight.push_back(kv.first);
        }
        mpdev_submit_work(kv.first, host_worker, t);
    }
    return mppipe_wait(p);
}
