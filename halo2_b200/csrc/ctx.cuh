// Host-side context shared by the translation units of the C ABI (capi_*.cu): error reporting, device buffers, the per-device
// Context, launch / profiling helpers.  The library is split into several TUs so that they compile in parallel and a change
// to one kernel family rebuilds one of them; every kernel is a template (or inline) in a header, so each TU instantiates
// what it launches.  No torch types.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/halo2_b200.h"
#define H2_MAX_UPLOAD_CHUNKS 4
#define H2_MAX_DEVICES 16
#include "curve.cuh"

using namespace h2;

// ------------------------------------------------------------------------------------------------
// errors, context
// ------------------------------------------------------------------------------------------------
int fail(const std::string &m);          // sets the calling thread's last error, returns 1
const std::string &last_error_string();  // the calling thread's last error (worker threads hand theirs to the caller)
#define CU(expr)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess) return fail(std::string(#expr) + ": " + cudaGetErrorString(e_));          \
    } while (0)

// Cached CUDA graphs (fixed-base MSMs) hold raw pointers into the library's scratch pools and window tables: every
// (re)allocation or release of a TRACKED buffer bumps the generation and invalidates them.  Buffers a graph can only see
// through its key (caller polynomials, IPA session vectors: the scalars / out pointers are part of the key) are untracked --
// allocating a ResidentPoly between two commits must not throw the commit graphs away.
extern std::atomic<uint64_t> g_alloc_gen;
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool tracked = true;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (tracked) g_alloc_gen++;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return fail(std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e)); }
        cap = want;
        return 0;
    }
    void release() { if (p) { cudaFree(p); if (tracked) g_alloc_gen++; } p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};


struct TwiddleEntry { int field; uint32_t log_n; uint8_t omega[32]; DevBuf buf; uint64_t stamp; };
struct BaseSet {
    int curve; size_t n; DevBuf buf;
    DevBuf table; uint32_t c = 0, W = 0;   // W x n window shifts 2^(c w) G_i (bucket method over one shared bucket set)
    DevBuf dtable;                         // 32 x 128 x n digit multiples m 2^(8 w) G_i (fixedbase.cuh: direct sum, small sets)
};

struct PolyBuf { int field; size_t len; DevBuf buf; PolyBuf() { buf.tracked = false; } };   // device-resident polynomial, Montgomery form, len + 1 slots
struct IpaSession {
    uint64_t bases; uint32_t k, round; int folded; DevBuf p, b, s, scal, out;
    IpaSession() { p.tracked = b.tracked = s.tracked = scal.tracked = out.tracked = false; }
};

// A fixed-base MSM over resident bases is ~25 small launches whose parameters repeat call after call (same table, same
// scratch, same sizes): the second call with a given key is captured into a CUDA graph, later ones replay it.
struct MsmGraph {
    const void *scalars, *bases, *out;
    size_t n; uint64_t stride, gen;
    uint32_t c, sets; int scalars_mont, out_canonical;
    uint32_t fast = 0;
    uint32_t seen = 0; uint64_t launches = 0, stamp = 0;
    cudaGraphExec_t exec = nullptr;
};

// Pinned staging ring for transfers from / to PAGEABLE caller memory (a Rust Vec, a numpy array): a plain cudaMemcpyAsync
// from pageable memory is staged by the driver through one thread and runs at a fraction of the link rate, and it blocks
// the caller so nothing overlaps.  Here a small pool of host threads copies slot-sized pieces into pinned slots while the
// DMA engine drains the previous ones (capi_core.cu: upload_async / download_sync).
struct StageRing {
    enum { SLOTS = 4 };
    uint8_t *slot[SLOTS] = {};
    size_t slot_bytes = 0;
    cudaEvent_t done[SLOTS] = {};
    bool busy[SLOTS] = {};
    uint32_t next = 0;
    int ensure();
    void destroy();
};

struct Context {
    bool ready = false;
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // uploads that may overlap compute (bases of a one-shot MSM)
    cudaEvent_t ev_scalars_up = nullptr, ev_bases_up[H2_MAX_UPLOAD_CHUNKS] = {}, ev_scal_up[H2_MAX_UPLOAD_CHUNKS] = {};
    uint32_t chunk_min_log = 19;             // one-shot MSMs of >= 2^19 points upload their bases in chunks
    cudaEvent_t last_use = nullptr;
    bool have_last = false;
    uint32_t window_override = 0;
    const uint32_t *last_flags = nullptr;    // device flags of the most recent MSM (test hook; the fast fixed-base pass's validity check)
    uint32_t fast_on = 1;                    // fixed-base passes over resident tables first run WITHOUT the fallback kernels (exact sort: histogram, 3 scan
                                             // kernels, scatter; partial merges: 3 kernels) and their two memsets -- 10 of ~27 graph nodes that do nothing on
                                             // ordinary inputs; the two device flags come back with the result and a set flag re-runs the full pass
    uint64_t natural_max_buckets = 1ull << 14;   // fast passes with up to this many buckets take them in index order (MsmPlan::natural).  Measured at k = 14
                                             // (one lane pair per bucket): 2^14 buckets (a single commit) 0.379 -> 0.356 ms; 2^15 (an IPA round's two sets, 1.15 waves)
                                             // 4.60 -> 4.74 ms per opening; 2^16 (4 batched commits) 0.71 -> 0.80 ms: only a pass that is resident at once gains
    bool fast_now = false;                   // ... the pass being issued is such a fast one
    bool last_fast = false;                  // ... the most recent pass was
    uint32_t *h_flags = nullptr;             // pinned host copy of the two flags
    uint32_t sort_bins = 1;                  // single-pass binned sort (0: always the exact two-pass sort)
    uint32_t glv_on = 1;                     // GLV endomorphism split for one-shot / table-less MSMs
    uint32_t accum_ways = 12;                // lanes per work item in the small-problem accumulation: 12 (default) / 14 = 2 / 4 INDEPENDENT lanes, each adding every
                                             // 2nd / 4th reference serially, partial sums folded by shuffles (msm_accum0_split_kernel); 0 = a cooperating PAIR
                                             // (10 lane-multiplies per addition, 5 levels); 1 / 2 / 4 = quads.  A cooperative addition is no shorter than a serial
                                             // one in practice (3.3-4.4 us against 3.6 us), so cutting the longest bucket's chain in two is what helps.  Measured
                                             // at k = 14, c = 15, same box: commit 0.367 (2 lanes) / 0.411 (pair) / 0.417 ms (4 lanes), 4 batched commits
                                             // 0.754 / 0.792 / 0.816 ms, IPA opening 4.58 / 4.58 / 4.97 ms; quads earlier: 0.407 ms, 5.68 ms
    uint32_t poly_cta = 0;                   // 1: eval_polynomial / kate_division of polynomials up to 2^16 coefficients in one CTA each; 0 (default): the level
                                             // tree.  Measured in the proof replay: k = 14 evaluations 0.73 (tree) vs 0.78 ms, k = 16 0.86 vs 1.36 ms, quotients
                                             // 0.40 vs 0.78 ms -- one CTA's 16-64-step serial slices lose to three launches that fill the machine
    uint64_t small_accum_refs = 1ull << 20;  // MSMs with up to this many references accumulate with cooperating lanes (accum_ways); larger ones with a thread per item
    uint32_t ntt_tma = 0;                    // NTT passes: 1 = bulk-copy (TMA) persistent kernel where it applies, 0 = classic kernel.  Measured on
                                             // B200 (profiles/r2e_*): 2^20 0.27-0.30 ms vs 0.215 ms, 2^24 4.77 vs 3.67 ms -- the pass is bound by the
                                             // integer pipes (fmaheavy 55-61 %, ALU 54 %, issue 51 %, top stall `wait`), not by its memory phases, and
                                             // the double-buffered tiles cost occupancy (4 CTAs/SM, 2 in the last pass): opt-in, default off
    uint32_t ba_variant = 3;                 // kernel variant of the rounds (tuning): 0 / 1 = gather chunks of 4 pairs at 4 / 5 CTAs per SM, 2 / 3 = chunks of 2
    uint32_t ba_rounds = 0, ba_target = 32;  // batched-affine halving rounds ahead of the XYZZ accumulation of large one-shot MSMs and the pairs per thread
                                             // one inversion is shared by (h2_test_set_batched_affine).  OFF by default -- measured on B200 at 2^20
                                             // (profiles/r2k_ba_sweep.txt): best setting (1 round, 32 pairs) 3.62 ms vs 3.61 ms without; 2 / 3 rounds
                                             // 3.84 / 4.04 ms.  An affine addition is 6 multiplies instead of 10 but 2 520 instructions against 2 640
                                             // (field add/sub, the inversion's share, call marshalling, local-memory products), and its dependent
                                             // reference -> point gathers leave the kernel latency-bound at 14-20 warps per SM
    uint32_t ecfft_quad = 1;                 // EC-FFT butterfly form: 1 = by size (default), 0 = one thread each, 2 = quads (test hook)
    // MSM scratch
    DevBuf scal_in, bases_in, bases_phi, glv_parts, scal_canon, counts, cursor, refs, size_hist, items, bucket_sum, pkey, pstart, pend, ppt, ra_t, ra_e, r0, r1,
        wsum, scan_blocks, result, misc, ba_lv[3];
    // NTT scratch
    DevBuf ntt_io, ntt_out, ntt_work, pow2;
    // EC-FFT / batch-normalise scratch: XYZZ work array (128 B per point), staging for the host forms
    DevBuf ec_work, ec_io, ec_out;
    DevBuf fb_a, fb_b;                       // partial sums of the direct-sum fixed-base MSM (ping-pong)
    DevBuf multi_parts;                      // primary device: the per-GPU partial results of a multi-GPU MSM (peer-written)
    StageRing stage;
    DevBuf ast_code, ast_consts;             // asteval.cuh: the postfix program and its constants
    DevBuf po_lvl, po_q, po_pts, po_ptrs;    // polyops.cuh: level arrays, kate carries, per-level points, pointer arrays
    DevBuf lk_keys, lk_left, lk_u32;         // lookup.cuh: sorted canonical keys (input | table), leftover table values, flag / scan arrays
    std::vector<TwiddleEntry *> twiddles;
    uint64_t tw_stamp = 0;
    std::map<uint64_t, BaseSet *> bases;
    std::map<uint64_t, IpaSession *> ipa;
    std::map<uint64_t, PolyBuf *> polys;
    std::vector<MsmGraph> graphs;
    uint64_t graph_stamp = 0;
    uint32_t graphs_on = 1;
    std::vector<PolyBuf *> poly_pool;        // freed resident polynomials keep their buffers for the next h2_poly_alloc of that size
    size_t poly_pool_bytes = 0;
    std::vector<IpaSession *> ipa_pool;      // finished sessions keep their buffers for the next proof (no cudaMalloc per opening)
    uint64_t next_handle = 1;
};
// One Context per CUDA device.  API functions work on the PRIMARY context (the device h2_init bound); the multi-GPU
// entry points (h2_multi_*) run one worker thread per device, each of which points its thread-local g_cur at its device's
// context while the calling thread holds g_mu -- so all the single-GPU code below runs unchanged, concurrently, on every
// device.
extern Context g_ctxs[H2_MAX_DEVICES];
extern Context *g_primary;
extern thread_local Context *g_cur;
static inline Context &cur_ctx() { return *(g_cur ? g_cur : g_primary); }
#define g_ctx (cur_ctx())
extern std::mutex g_mu;
// optional per-kernel timing (bench.py's roofline leg): event pairs recorded on the launch stream
struct ProfSpan { int kind; cudaEvent_t e0, e1; };
extern bool g_prof_on;
extern std::vector<ProfSpan> g_prof;
enum { PROF_MSM_ACCUM0 = 0, PROF_NTT_PASS = 1, PROF_KINDS = 2 };
void prof_begin(int kind, cudaStream_t s);
void prof_end(cudaStream_t s);
extern std::atomic<uint64_t> g_launches;

#define LAUNCH(kernel, grid, block, smem, stream, ...)                                                   \
    do {                                                                                                 \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                      \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                              \
        cudaError_t e_ = cudaGetLastError();                                                             \
        if (e_ != cudaSuccess) return fail(std::string(#kernel) + " launch: " + cudaGetErrorString(e_)); \
    } while (0)

// Host -> device copy on stream `s` that keeps the link busy whatever the caller's memory is: pinned / registered memory
// goes straight to cudaMemcpyAsync (the source must then stay valid until the stream has run it -- every caller synchronises
// before returning); pageable memory goes through the context's pinned ring and has been fully READ when this returns.
int upload_async(void *d_dst, const void *h_src, size_t bytes, cudaStream_t s);
// Device -> host copy ordered behind the work already on `s`; returns when the data is in h_dst.
int download_sync(void *h_dst, const void *d_src, size_t bytes, cudaStream_t s);
void h2_set_staging(int on);             // test / bench hook: 0 = always plain cudaMemcpyAsync
int require_ready();
int scratch_acquire(cudaStream_t s);     // make `s` wait for whatever last used the shared scratch
int scratch_release(cudaStream_t s);
static inline uint32_t blocks_for(uint64_t n, uint32_t bs) { return (uint32_t)((n + bs - 1) / bs); }

template <class P> static fe host_to_mont(const void *bytes, int repr) {
    fe x;
    memcpy(x.v, bytes, 32);
    return repr == H2_REPR_MONTGOMERY ? x : fe_to_mont<P>(x);
}
static inline PolyBuf *find_poly(uint64_t h) {
    auto it = g_ctx.polys.find(h);
    return it == g_ctx.polys.end() ? nullptr : it->second;
}
// what a fixed-base MSM over `b` runs on: the digit-multiples table (mode 2) when there is one, else the window table (mode 1)
#define H2_FB_BITS_CTX 8u
static inline const affine *fixed_table(const BaseSet *b, uint32_t *c, uint32_t *mode) {
    if (b->dtable.p) { *c = H2_FB_BITS_CTX; *mode = 2; return b->dtable.as<affine>(); }
    *c = b->c; *mode = 1;
    return b->table.as<affine>();
}

// ---- functions one TU defines and others call -------------------------------------------------------------------
// Arrival of a one-shot MSM's inputs in k chunks: events on the copy stream -- bases / scalars of chunk j have landed.
// When the inputs are staged from pageable memory an uploader thread records the events while the calling thread issues
// the kernels: `recorded` (2 j + 1 after the scalars of chunk j, 2 j + 2 after its bases) tells the issuer that an event
// HAS been recorded and may be waited on; `failed` aborts the issue.
struct BasesChunks {
    uint32_t k = 0;
    cudaEvent_t ev[H2_MAX_UPLOAD_CHUNKS], ev_scal[H2_MAX_UPLOAD_CHUNKS];
    std::atomic<uint32_t> *recorded = nullptr;
    std::atomic<int> *failed = nullptr;
    int wait_recorded(uint32_t want) const {
        if (!recorded) return 0;
        while (recorded->load(std::memory_order_acquire) < want) {
            if (failed && failed->load()) return 1;
            std::this_thread::yield();
        }
        return 0;
    }
};
// capi_msm.cu
int msm_dispatch(int curve, const fe *d_scalars, int scalars_mont, const affine *d_bases, size_t n, uint32_t c,
                 jacobian *d_out, int out_canonical, cudaStream_t s, uint32_t fixed = 0, uint64_t stride = 0,
                 const BasesChunks *bc = nullptr, uint32_t sets = 1);
int convert_points(int curve, affine *d, size_t n, int to_mont, cudaStream_t s);
// capi_ntt.cu
int get_twiddles_any(int field, const fe &omega_mont, uint32_t log_n, cudaStream_t s, const fe **out);
int convert_field(int field, fe *d, size_t n, int to_mont, cudaStream_t s);
