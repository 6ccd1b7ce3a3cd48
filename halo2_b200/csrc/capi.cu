// C ABI of the engine (include/halo2_b200.h): context, scratch management, kernel launch
// sequences for the MSM and NTT pipelines, host<->device staging.  No torch types.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/halo2_b200.h"
#define H2_MAX_UPLOAD_CHUNKS 4
#define H2_MSM_QUAD_ACCUM_REFS (1ull << 20)   // up to this many references the accumulation runs one quad per work item
#include "msm.cuh"
#include "ipa.cuh"
#include "ntt.cuh"
#include "ecfft.cuh"
#include "fixedbase.cuh"
#include "codec.cuh"
#include "h2c.cuh"
#include "polyops.cuh"
#include "asteval.cuh"

using namespace h2;

// ------------------------------------------------------------------------------------------------
// errors, context
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(const std::string &m) { g_err = m; return 1; }
#define CU(expr)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess) return fail(std::string(#expr) + ": " + cudaGetErrorString(e_));          \
    } while (0)

static uint64_t g_alloc_gen = 0;   // bumped by every (re)allocation or release: cached CUDA graphs hold raw pointers
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        g_alloc_gen++;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return fail(std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e)); }
        cap = want;
        return 0;
    }
    void release() { if (p) { cudaFree(p); g_alloc_gen++; } p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct TwiddleEntry { int field; uint32_t log_n; uint8_t omega[32]; DevBuf buf; uint64_t stamp; };
struct BaseSet {
    int curve; size_t n; DevBuf buf;
    DevBuf table; uint32_t c = 0, W = 0;   // W x n window shifts 2^(c w) G_i (bucket method over one shared bucket set)
    DevBuf dtable;                         // 32 x 128 x n digit multiples m 2^(8 w) G_i (fixedbase.cuh: direct sum, small sets)
};

struct PolyBuf { int field; size_t len; DevBuf buf; };   // device-resident polynomial, Montgomery form, len + 1 slots
struct IpaSession { uint64_t bases; uint32_t k, round; int folded; DevBuf p, b, s, scal, out; };

// A fixed-base MSM over resident bases is ~25 small launches whose parameters repeat call after call (same table, same
// scratch, same sizes): the second call with a given key is captured into a CUDA graph, later ones replay it.
struct MsmGraph {
    const void *scalars, *bases, *out;
    size_t n; uint64_t stride, gen;
    uint32_t c, sets; int scalars_mont, out_canonical;
    uint32_t seen = 0; uint64_t launches = 0, stamp = 0;
    cudaGraphExec_t exec = nullptr;
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
    const uint32_t *last_flags = nullptr;    // device flags of the most recent MSM (test hook)
    uint32_t sort_bins = 1;                  // single-pass binned sort (0: always the exact two-pass sort)
    uint32_t glv_on = 1;                     // GLV endomorphism split for one-shot / table-less MSMs
    uint32_t accum_ways = 1;                 // quads per work item in the small-problem accumulation (1, 2, 4; test hook).  Measured at
                                             // k = 14, c = 15: commit 0.360 / 0.329 / 0.361 ms, IPA opening 5.7 / 6.1 / 7.6 ms -- the accumulation is
                                             // bound by lane-multiplies (a quad addition occupies 16 slots for 10 products), not by its chains
    uint32_t ecfft_quad = 1;                 // EC-FFT butterfly form: 1 = by size (default), 0 = one thread each, 2 = quads (test hook)
    // MSM scratch
    DevBuf scal_in, bases_in, bases_phi, glv_parts, scal_canon, counts, cursor, refs, size_hist, items, bucket_sum, pkey, pstart, pend, ppt, ra_t, ra_e, r0, r1,
        wsum, scan_blocks, result, misc;
    // NTT scratch
    DevBuf ntt_io, ntt_out, ntt_work, pow2;
    // EC-FFT / batch-normalise scratch: XYZZ work array (128 B per point), staging for the host forms
    DevBuf ec_work, ec_io, ec_out;
    DevBuf fb_a, fb_b;                       // partial sums of the direct-sum fixed-base MSM (ping-pong)
    DevBuf ast_code, ast_consts;             // asteval.cuh: the postfix program and its constants
    DevBuf po_lvl, po_q, po_pts, po_ptrs;    // polyops.cuh: level arrays, kate carries, per-level points, pointer arrays
    std::vector<TwiddleEntry *> twiddles;
    uint64_t tw_stamp = 0;
    std::map<uint64_t, BaseSet *> bases;
    std::map<uint64_t, IpaSession *> ipa;
    std::map<uint64_t, PolyBuf *> polys;
    std::vector<MsmGraph> graphs;
    uint64_t graph_stamp = 0;
    uint32_t graphs_on = 1;
    std::vector<IpaSession *> ipa_pool;      // finished sessions keep their buffers for the next proof (no cudaMalloc per opening)
    uint64_t next_handle = 1;
};
static Context g_ctx;
static std::mutex g_mu;
// optional per-kernel timing (bench.py's roofline leg): event pairs recorded on the launch stream
struct ProfSpan { int kind; cudaEvent_t e0, e1; };
static bool g_prof_on = false;
static std::vector<ProfSpan> g_prof;
enum { PROF_MSM_ACCUM0 = 0, PROF_NTT_PASS = 1, PROF_KINDS = 2 };
static void prof_begin(int kind, cudaStream_t s) {
    if (!g_prof_on) return;
    ProfSpan sp; sp.kind = kind;
    cudaEventCreate(&sp.e0); cudaEventCreate(&sp.e1);
    cudaEventRecord(sp.e0, s);
    g_prof.push_back(sp);
}
static void prof_end(cudaStream_t s) {
    if (!g_prof_on || g_prof.empty()) return;
    cudaEventRecord(g_prof.back().e1, s);
}
static std::atomic<uint64_t> g_launches{0};

#define LAUNCH(kernel, grid, block, smem, stream, ...)                                                   \
    do {                                                                                                 \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                      \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                              \
        cudaError_t e_ = cudaGetLastError();                                                             \
        if (e_ != cudaSuccess) return fail(std::string(#kernel) + " launch: " + cudaGetErrorString(e_)); \
    } while (0)

static int require_ready() {
    if (!g_ctx.ready) return fail("h2_init has not been called (or failed): no CUDA device bound; there is no CPU fallback");
    CU(cudaSetDevice(g_ctx.device));
    return 0;
}
// make `s` wait for whatever last used the shared scratch
static int scratch_acquire(cudaStream_t s) {
    if (g_ctx.have_last) CU(cudaStreamWaitEvent(s, g_ctx.last_use, 0));
    return 0;
}
static int scratch_release(cudaStream_t s) {
    CU(cudaEventRecord(g_ctx.last_use, s));
    g_ctx.have_last = true;
    return 0;
}

extern "C" const char *h2_last_error(void) { return g_err.c_str(); }
extern "C" uint32_t h2_abi_version(void) { return 1; }
extern "C" uint64_t h2_launch_count(void) { return g_launches.load(); }
extern "C" int h2_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
extern "C" int h2_init(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_ctx.ready && g_ctx.device == device) return 0;
    if (g_ctx.ready) return fail("h2_init: already bound to another device (one process per GPU)");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(std::string("h2_init: no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail("h2_init: device index out of range");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail("h2_init: this library is built for sm_100a (B200) only");
    CU(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&g_ctx.last_use, cudaEventDisableTiming));
    CU(cudaStreamCreateWithFlags(&g_ctx.copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&g_ctx.ev_scalars_up, cudaEventDisableTiming));
    for (int j = 0; j < H2_MAX_UPLOAD_CHUNKS; j++) {
        CU(cudaEventCreateWithFlags(&g_ctx.ev_bases_up[j], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&g_ctx.ev_scal_up[j], cudaEventDisableTiming));
    }
    g_ctx.device = device;
    g_ctx.ready = true;
    return 0;
}
extern "C" int h2_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ctx.ready) return 0;
    cudaSetDevice(g_ctx.device);
    cudaDeviceSynchronize();
    DevBuf *all[] = {&g_ctx.scal_in, &g_ctx.bases_in, &g_ctx.bases_phi, &g_ctx.glv_parts, &g_ctx.scal_canon, &g_ctx.counts, &g_ctx.cursor, &g_ctx.refs, &g_ctx.size_hist,
                     &g_ctx.items, &g_ctx.bucket_sum, &g_ctx.pkey, &g_ctx.pstart, &g_ctx.pend, &g_ctx.ppt, &g_ctx.ra_t, &g_ctx.ra_e,
                     &g_ctx.r0, &g_ctx.r1, &g_ctx.wsum, &g_ctx.scan_blocks, &g_ctx.result, &g_ctx.misc, &g_ctx.ntt_io, &g_ctx.ntt_out,
                     &g_ctx.ntt_work, &g_ctx.pow2, &g_ctx.ec_work, &g_ctx.ec_io, &g_ctx.ec_out, &g_ctx.fb_a, &g_ctx.fb_b, &g_ctx.po_lvl, &g_ctx.po_q, &g_ctx.po_pts, &g_ctx.po_ptrs, &g_ctx.ast_code, &g_ctx.ast_consts};
    for (DevBuf *b : all) b->release();
    for (auto *t : g_ctx.twiddles) { t->buf.release(); delete t; }
    g_ctx.twiddles.clear();
    for (auto &kv : g_ctx.bases) { kv.second->buf.release(); kv.second->table.release(); kv.second->dtable.release(); delete kv.second; }
    g_ctx.bases.clear();
    for (auto &kv : g_ctx.ipa) { IpaSession *q = kv.second; q->p.release(); q->b.release(); q->s.release(); q->scal.release(); q->out.release(); delete q; }
    g_ctx.ipa.clear();
    for (auto &kv : g_ctx.polys) { kv.second->buf.release(); delete kv.second; }
    g_ctx.polys.clear();
    for (auto &ge : g_ctx.graphs) if (ge.exec) cudaGraphExecDestroy(ge.exec);
    g_ctx.graphs.clear();
    for (IpaSession *q : g_ctx.ipa_pool) { q->p.release(); q->b.release(); q->s.release(); q->scal.release(); q->out.release(); delete q; }
    g_ctx.ipa_pool.clear();
    cudaEventDestroy(g_ctx.ev_scalars_up);
    for (int j = 0; j < H2_MAX_UPLOAD_CHUNKS; j++) { cudaEventDestroy(g_ctx.ev_bases_up[j]); cudaEventDestroy(g_ctx.ev_scal_up[j]); }
    cudaStreamDestroy(g_ctx.copy_stream);
    cudaEventDestroy(g_ctx.last_use);
    cudaStreamDestroy(g_ctx.stream);
    g_ctx = Context();
    return 0;
}
extern "C" int h2_set_glv(int on) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_ctx.glv_on = on ? 1u : 0u;
    return 0;
}
extern "C" int h2_set_sort_mode(int exact_only) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_ctx.sort_bins = exact_only ? 0u : 1u;
    return 0;
}
// test hook: flags of the most recent MSM -- bit 0: some bucket was split into several work items, bit 1: the exact
// sort ran (bin overflow, or no bins).  Synchronises the device.
extern "C" int h2_test_last_msm_flags(uint32_t *out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    if (!g_ctx.last_flags) return fail("h2_test_last_msm_flags: no MSM has run");
    uint32_t f[2];
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(f, g_ctx.last_flags, sizeof f, cudaMemcpyDeviceToHost));
    *out = (f[0] ? 1u : 0u) | (f[1] ? 2u : 0u);
    return 0;
}
// test hook: CUDA-graph replay of fixed-base MSMs on / off
extern "C" int h2_test_set_graphs(int on) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_ctx.graphs_on = on ? 1u : 0u;
    return 0;
}
// test hook: quads per work item of the small-problem accumulation (1, 2 or 4).  Invalidates nothing: graphs are keyed by
// their parameters only, so flip it before the first fixed-base MSM of a base set or with graphs off.
extern "C" int h2_test_set_accum_ways(uint32_t ways) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (ways != 1 && ways != 2 && ways != 4) return fail("h2_test_set_accum_ways: 1, 2 or 4");
    g_ctx.accum_ways = ways;
    for (auto &ge : g_ctx.graphs) if (ge.exec) { cudaGraphExecDestroy(ge.exec); ge.exec = nullptr; ge.seen = 0; }
    return 0;
}
// test hook: EC-FFT butterfly form -- 1: quads of lanes, 0: one thread each, -1: by size (the default)
extern "C" int h2_test_set_ecfft_quad(int on) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_ctx.ecfft_quad = on < 0 ? 1u : on ? 2u : 0u;   // -1: by size (default), 0: thread form, 1: quad form
    return 0;
}
// test hook: one-shot MSMs (h2_msm) of >= 2^log2_n points upload their bases in chunks (default 19)
extern "C" int h2_test_set_chunk_threshold(uint32_t log2_n) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (log2_n > 40) return fail("h2_test_set_chunk_threshold: log2_n > 40");
    g_ctx.chunk_min_log = log2_n;
    return 0;
}
extern "C" int h2_set_window_bits(uint32_t c) {
    if (c > 24) return fail("h2_set_window_bits: c must be <= 24");
    std::lock_guard<std::mutex> lk(g_mu);
    g_ctx.window_override = c;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// small kernels: conversions, generators, self-tests
// ------------------------------------------------------------------------------------------------
template <class P> __global__ void convert_kernel(fe *a, uint64_t n, int to_mont) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe x = fe_load(a + i);
    fe_store(a + i, to_mont ? fe_to_mont<P>(x) : fe_from_mont<P>(x));
}
// affine points: identity (0,0) maps to itself under both conversions
template <class P> __global__ void convert_points_kernel(affine *a, uint64_t n, int to_mont) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    affine p = ld_affine(a + i);
    if (to_mont) { p.x = fe_to_mont<P>(p.x); p.y = fe_to_mont<P>(p.y); }
    else { p.x = fe_from_mont<P>(p.x); p.y = fe_from_mont<P>(p.y); }
    st_affine(a + i, p);
}
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
template <class P> __device__ affine xyzz_to_affine_dev(const xyzz &p) {
    affine r;
    if (xyzz_is_identity(p)) { r.x = fe_zero(); r.y = fe_zero(); return r; }
    fe t = fe_inv<P>(fe_mul<P>(p.zz, p.zzz));
    r.x = fe_mul<P>(p.x, fe_mul<P>(t, p.zzz));   // X / ZZ
    r.y = fe_mul<P>(p.y, fe_mul<P>(t, p.zz));    // Y / ZZZ
    return r;
}
template <class P> __global__ void __launch_bounds__(128) gen_points_kernel(affine *out, uint64_t seed, uint64_t first, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s = splitmix64(seed ^ splitmix64(first + i)) | 1ULL;
    affine g;
    g.x = fe_neg<P>(fe_one<P>());
    g.y = fe_dbl<P>(fe_one<P>());
    xyzz acc = xyzz_identity();
    for (int b = 63; b >= 0; b--) {
        xyzz_double<P>(acc);
        if ((s >> b) & 1ULL) xyzz_add_mixed<P>(acc, g);
    }
    st_affine(out + i, xyzz_to_affine_dev<P>(acc));
}
template <class P> __global__ void test_field_kernel(const fe *a, const fe *b, fe *out, uint64_t n, int op) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe x = fe_to_mont<P>(fe_load(a + i)), y = fe_to_mont<P>(fe_load(b + i)), r;
    switch (op) {
    case 0: r = fe_add<P>(x, y); break;
    case 1: r = fe_sub<P>(x, y); break;
    case 2: r = fe_mul<P>(x, y); break;
    case 3: r = fe_inv<P>(x); break;
    default: r = fe_sqr<P>(x); break;
    }
    fe_store(out + i, fe_from_mont<P>(r));
}
template <class P> __global__ void __launch_bounds__(64) test_curve_kernel(const affine *a, const affine *b, affine *out, uint64_t n, int op) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    affine pa = ld_affine(a + i), pb = ld_affine(b + i);
    if (!affine_is_identity(pa)) { pa.x = fe_to_mont<P>(pa.x); pa.y = fe_to_mont<P>(pa.y); }
    xyzz r = xyzz_from_affine<P>(pa);
    if (op == 0) {
        if (!affine_is_identity(pb)) { pb.x = fe_to_mont<P>(pb.x); pb.y = fe_to_mont<P>(pb.y); }
        xyzz t = r;
        xyzz_add_mixed<P>(r, pb);                         // mixed path
        xyzz full = xyzz_from_affine<P>(pb);
        xyzz_add<P>(t, full);                             // full-add path must agree
        affine r1 = xyzz_to_affine_dev<P>(r), r2 = xyzz_to_affine_dev<P>(t);
        if (!(fe_eq(r1.x, r2.x) && fe_eq(r1.y, r2.y))) { r1.x = fe_one<P>(); r1.y = fe_zero(); }   // poison
        r1.x = fe_from_mont<P>(r1.x); r1.y = fe_from_mont<P>(r1.y);
        st_affine(out + i, r1);
        return;
    } else if (op == 1) {
        xyzz_double<P>(r);
    } else {
        uint32_t k[8];
        for (int j = 0; j < 8; j++) k[j] = pb.x.v[j];
        r = xyzz_scalar_mul<P>(pa, k);
    }
    affine o = xyzz_to_affine_dev<P>(r);
    o.x = fe_from_mont<P>(o.x); o.y = fe_from_mont<P>(o.y);
    st_affine(out + i, o);
}
// throughput microbenchmark: 4 independent dependent-chains per thread
template <class P, bool SQR> __global__ void bench_mul_kernel(fe *io, uint32_t iters) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    fe a = fe_load(io + 4 * i), b = fe_load(io + 4 * i + 1), c = fe_load(io + 4 * i + 2), d = fe_load(io + 4 * i + 3);
    for (uint32_t k = 0; k < iters; k++) {
        if (SQR) { a = fe_sqr<P>(a); b = fe_sqr<P>(b); c = fe_sqr<P>(c); d = fe_sqr<P>(d); }
        else { a = fe_mul<P>(a, b); b = fe_mul<P>(b, c); c = fe_mul<P>(c, d); d = fe_mul<P>(d, a); }
    }
    fe_store(io + 4 * i, a); fe_store(io + 4 * i + 1, b); fe_store(io + 4 * i + 2, c); fe_store(io + 4 * i + 3, d);
}
// single-warp latency microbenchmark of the serial building blocks (tails of the MSM)
template <class P> __global__ void bench_latency_kernel(fe *io, uint32_t iters, int mode) {
    uint32_t i = threadIdx.x;
    fe a = fe_load(io + 4 * i), b = fe_load(io + 4 * i + 1), c = fe_load(io + 4 * i + 2), d = fe_load(io + 4 * i + 3);
    a.v[7] &= 0x3fffffffu; b.v[7] &= 0x3fffffffu; c.v[7] &= 0x3fffffffu; d.v[7] &= 0x3fffffffu;
    if (mode == 0) { for (uint32_t k = 0; k < iters; k++) a = fe_mul<P>(a, b); }
    else if (mode == 1) { for (uint32_t k = 0; k < iters; k++) { a = fe_mul<P>(a, b); c = fe_mul<P>(c, d); } }
    else if (mode == 2) { for (uint32_t k = 0; k < iters; k++) { a = fe_mul<P>(a, b); b = fe_mul<P>(b, c); c = fe_mul<P>(c, d); d = fe_mul<P>(d, a); } }
    else if (mode == 6) { for (uint32_t k = 0; k < iters; k++) { fe_mul2<P>(a, a, b, c, c, d); } }
    else {
        affine g; g.x = fe_neg<P>(fe_one<P>()); g.y = fe_dbl<P>(fe_one<P>());
        xyzz acc = xyzz_double_affine<P>(g), other = acc; xyzz_double<P>(other);
        if (mode == 3) { for (uint32_t k = 0; k < iters; k++) xyzz_double<P>(acc); }
        else if (mode == 4) { for (uint32_t k = 0; k < iters; k++) xyzz_add<P>(acc, other); }
        else { for (uint32_t k = 0; k < iters; k++) xyzz_add_mixed<P>(acc, g); }
        a = acc.x; b = acc.y; c = acc.zz; d = acc.zzz;
    }
    fe_store(io + 4 * i, a); fe_store(io + 4 * i + 1, b); fe_store(io + 4 * i + 2, c); fe_store(io + 4 * i + 3, d);
}
template <class P> __global__ void point_sum_kernel(const jacobian *pts, uint32_t g, int canonical, jacobian *out) {
    if (threadIdx.x || blockIdx.x) return;
    xyzz acc = xyzz_identity();
    for (uint32_t i = 0; i < g; i++) {
        jacobian j;
        j.x = fe_load(&pts[i].x); j.y = fe_load(&pts[i].y); j.z = fe_load(&pts[i].z);
        if (canonical) { j.x = fe_to_mont<P>(j.x); j.y = fe_to_mont<P>(j.y); j.z = fe_to_mont<P>(j.z); }
        xyzz t;
        if (fe_is_zero(j.z)) t = xyzz_identity();
        else { t.x = j.x; t.y = j.y; t.zz = fe_sqr<P>(j.z); t.zzz = fe_mul<P>(t.zz, j.z); }
        xyzz_add<P>(acc, t);
    }
    jacobian r = xyzz_to_jacobian<P>(acc);
    if (canonical) { r.x = fe_from_mont<P>(r.x); r.y = fe_from_mont<P>(r.y); r.z = fe_from_mont<P>(r.z); }
    st_jacobian(out, r);
}

static inline uint32_t blocks_for(uint64_t n, uint32_t bs) { return (uint32_t)((n + bs - 1) / bs); }

// ------------------------------------------------------------------------------------------------
// MSM pipeline
// ------------------------------------------------------------------------------------------------
static int exclusive_scan_u32(uint32_t *d, uint64_t n, cudaStream_t s, const uint32_t *only_if = nullptr) {
    const uint64_t per_block = (uint64_t)H2_SCAN_BLOCK * H2_SCAN_ITEMS;
    uint32_t nb = (uint32_t)((n + per_block - 1) / per_block);
    if (g_ctx.scan_blocks.ensure((size_t)nb * 4 + 16)) return 1;
    uint32_t *bs = g_ctx.scan_blocks.as<uint32_t>();
    LAUNCH(scan_block_sums_kernel, nb, H2_SCAN_BLOCK, 0, s, d, n, bs, only_if);
    LAUNCH(scan_single_block_kernel, 1, H2_SCAN_BLOCK, 0, s, bs, nb, only_if);
    LAUNCH(scan_apply_kernel, nb, H2_SCAN_BLOCK, 0, s, d, n, bs, only_if);
    return 0;
}

// Arrival of the inputs of a one-shot MSM in `k` chunks (events on the copy stream): chunk j = points
// [chunk_first(n, k, j), chunk_first(n, k, j + 1)).  The chunks GROW: nothing can run before the first chunk has
// landed, so it is small (1/16 - 1/4 of the points), and the accumulation of chunk j hides the upload of the larger
// chunk j + 1 -- the link stays busy from t = 0 and the GPU from the first chunk's arrival.  (Equal chunks left the
// GPU idle for 1/k of the upload time: 2^20 pairs, 96 MiB at ~50 GB/s, 2 chunks: 0.95 of 4.57 ms.)
static inline size_t chunk_first(size_t n, uint32_t k, uint32_t j) {
    static const uint32_t cut[H2_MAX_UPLOAD_CHUNKS + 1][H2_MAX_UPLOAD_CHUNKS + 1] = {
        {0, 16, 16, 16, 16}, {0, 16, 16, 16, 16}, {0, 4, 16, 16, 16}, {0, 2, 8, 16, 16}, {0, 1, 4, 10, 16}};   // sixteenths
    if (k > H2_MAX_UPLOAD_CHUNKS) k = H2_MAX_UPLOAD_CHUNKS;
    if (j >= k) return n;
    return (size_t)((unsigned __int128)n * cut[k][j] / 16);
}
struct BasesChunks { uint32_t k = 0; cudaEvent_t ev[H2_MAX_UPLOAD_CHUNKS], ev_scal[H2_MAX_UPLOAD_CHUNKS]; };   // bases / scalars of chunk j have landed

// A fixed-base MSM over resident bases is launched with the same parameters call after call: the second call with a given
// key is captured into a CUDA graph, later ones replay it.
static int msm_issue_or_replay(const std::function<int()> &issue, bool graphable, const void *d_scalars, const void *d_bases, const void *d_out,
                               size_t n, uint64_t stride, uint32_t c, uint32_t sets, int scalars_mont, int out_canonical, cudaStream_t s) {
    Context &X = g_ctx;
    if (!(graphable && X.graphs_on && !g_prof_on)) return issue();
    MsmGraph *ge = nullptr;
    for (auto &e : X.graphs)
        if (e.scalars == d_scalars && e.bases == d_bases && e.out == d_out && e.n == n && e.stride == stride && e.c == c && e.sets == sets &&
            e.scalars_mont == scalars_mont && e.out_canonical == out_canonical) { ge = &e; break; }
    if (ge && ge->gen != g_alloc_gen) {   // some buffer moved since the capture
        if (ge->exec) cudaGraphExecDestroy(ge->exec);
        ge->exec = nullptr; ge->seen = 0; ge->gen = g_alloc_gen;
    }
    if (!ge) {
        if (X.graphs.size() >= 16) {   // evict the least recently used entry
            size_t v = 0;
            for (size_t i = 1; i < X.graphs.size(); i++) if (X.graphs[i].stamp < X.graphs[v].stamp) v = i;
            if (X.graphs[v].exec) cudaGraphExecDestroy(X.graphs[v].exec);
            X.graphs.erase(X.graphs.begin() + v);
        }
        MsmGraph e;
        e.scalars = d_scalars; e.bases = d_bases; e.out = d_out; e.n = n; e.stride = stride; e.gen = g_alloc_gen; e.c = c; e.sets = sets;
        e.scalars_mont = scalars_mont; e.out_canonical = out_canonical;
        X.graphs.push_back(e);
        ge = &X.graphs.back();
    }
    ge->stamp = ++X.graph_stamp;
    if (ge->exec) {
        CU(cudaGraphLaunch(ge->exec, s));
        g_launches.fetch_add(ge->launches, std::memory_order_relaxed);
        return 0;
    }
    if (ge->seen++ == 0) return issue();     // first sighting: run eagerly (the buffers may still be growing)
    const uint64_t l0 = g_launches.load();
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return issue(); }
    int rc = issue();
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(s, &graph);
    if (rc || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        ge->seen = 0;
        return rc ? rc : issue();            // capture refused: run eagerly
    }
    ge->launches = g_launches.load() - l0;
    ce = cudaGraphInstantiate(&ge->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { ge->exec = nullptr; cudaGetLastError(); return issue(); }
    CU(cudaGraphLaunch(ge->exec, s));
    return 0;
}

// fixed != 0: d_bases is a window table (stride points per window) built with window size c.
// bc != nullptr: the bases arrive chunk by chunk while this runs.  Each chunk is then sorted and accumulated on its own
// (own bins, work items and bucket sums) as soon as it has landed, and the bucket reduce adds the per-chunk bucket sums:
// the upload of all but the first chunk hides behind the accumulation.
template <class P, class PS>
static int msm_run(const fe *d_scalars, int scalars_mont, const affine *d_bases, size_t n, uint32_t c, uint32_t fixed, uint64_t stride,
                   jacobian *d_out, int out_canonical, cudaStream_t s, const BasesChunks *bc = nullptr, uint32_t sets = 1) {
    Context &X = g_ctx;
    if (n == 0) {   // empty sum = identity
        jacobian id;
        id.x = fe_zero(); id.y = out_canonical ? fe_zero() : fe_one<P>(); id.z = fe_zero();
        if (out_canonical) id.y.v[0] = 1;
        CU(cudaMemcpyAsync(d_out, &id, sizeof id, cudaMemcpyHostToDevice, s));
        CU(cudaStreamSynchronize(s));
        return 0;
    }
    std::function<int()> issue;
    if (fixed == 2) {   // direct sum over the digit-multiples table (fixedbase.cuh): accumulate + reduce tree
        FbPlan fp;
        fp.total = n; fp.sets = sets ? sets : 1u; fp.split = fb_split(n, fp.sets); fp.scalars_mont = scalars_mont ? 1u : 0u;
        const uint64_t count0 = n * fp.split;
        if (X.fb_a.ensure(fp.sets * count0 * sizeof(xyzz)) || X.fb_b.ensure(fp.sets * fb_ctas(count0, fb_fan(count0)) * sizeof(xyzz))) return 1;
        issue = [&X, fp, count0, d_scalars, d_bases, d_out, out_canonical, s]() -> int {
            auto k_acc = fb_accum_kernel<P, PS>;
            auto k_red = fb_reduce_kernel<P, PS>;
            xyzz *a = X.fb_a.as<xyzz>(), *b = X.fb_b.as<xyzz>();
            prof_begin(PROF_MSM_ACCUM0, s);
            LAUNCH(k_acc, blocks_for(fp.sets * count0, 128), 128, 0, s, fp, d_scalars, d_bases, a);
            prof_end(s);
            uint64_t count = count0, in_stride = count0;
            for (;;) {
                const uint32_t f = fb_fan(count);
                const uint64_t ctas = fb_ctas(count, f);
                LAUNCH(k_red, dim3((unsigned)ctas, fp.sets), 4 * H2_FB_QUADS, 0, s, (const xyzz *)a, count, in_stride, f, b, ctas,
                       ctas == 1 ? d_out : (jacobian *)nullptr, (uint32_t)out_canonical);
                if (ctas == 1) break;
                xyzz *t = a; a = b; b = t;
                count = ctas; in_stride = ctas;
            }
            return 0;
        };
        return msm_issue_or_replay(issue, !bc, d_scalars, d_bases, d_out, n, stride, H2_FB_BITS, fp.sets, scalars_mont, out_canonical, s);
    }
    const uint32_t glv = (!fixed && X.glv_on && n < (1ull << 30)) ? 1u : 0u;
    if (c == 0) c = X.window_override ? X.window_override : msm_default_window(n, glv);
    if (c > 24) return fail("msm: window bits > 24");
    const uint32_t K = (bc && !fixed && bc->k > 1) ? bc->k : 1u;
    const uint32_t force_cap = X.sort_bins ? 0u : H2_MSM_NO_BINS;
    MsmPlan p;                       // the whole problem: bucket reduce and window combine
    msm_make_plan(p, n, c, 0, 0, fixed, stride, glv, sets, force_cap);
    p.chunks = K;
    MsmPlan pk[H2_MAX_UPLOAD_CHUNKS];   // one chunk of points: sort, work items, accumulation
    size_t first[H2_MAX_UPLOAD_CHUNKS + 1];
    for (uint32_t j = 0; j <= K; j++) first[j] = chunk_first(n, K, j);
    uint64_t ref_space = 0, max_items = 0, part_total = 0;
    uint32_t t_max = 0;
    for (uint32_t j = 0; j < K; j++) {
        if (K == 1) pk[0] = p;
        else msm_make_plan(pk[j], first[j + 1] - first[j], c, 0, 0, fixed, stride, glv, sets, force_cap);
        if (pk[j].ref_space >= (1ull << 32)) return fail("msm: n * windows exceeds 2^32 references");
        ref_space = pk[j].ref_space > ref_space ? pk[j].ref_space : ref_space;
        max_items = pk[j].max_items > max_items ? pk[j].max_items : max_items;
        part_total = pk[j].part_total > part_total ? pk[j].part_total : part_total;
        t_max = pk[j].T > t_max ? pk[j].T : t_max;
    }
    if (glv && (X.bases_phi.ensure(n * sizeof(affine)) || X.glv_parts.ensure(n * 32))) return 1;
    if (fixed && (uint64_t)p.W * stride >= (1ull << 31)) return fail("msm: window table too large for 31-bit references");
    if (p.G >= (1ull << 32) || n >= (1ull << 31)) return fail("msm: n * windows exceeds 2^32 references");
    if (scalars_mont && X.scal_canon.ensure(n * p.sets * sizeof(fe))) return 1;
    const size_t small_words = 2 * (t_max + 2) + 8;   // size_hist (T + 2) | size_cursor (T + 1) | flags
    part_total += 1;
    if (X.counts.ensure(K * (p.G + 1) * 4) || X.cursor.ensure(K * 2 * p.G * 4) || X.refs.ensure(K * ref_space * 4) ||
        X.size_hist.ensure(K * small_words * 4) || X.items.ensure(K * max_items * sizeof(uint2)) ||
        X.bucket_sum.ensure(K * p.G * sizeof(xyzz)) || X.pkey.ensure(K * part_total * 4) || X.pstart.ensure(K * part_total * 4) ||
        X.pend.ensure(K * part_total * 4) || X.ppt.ensure(K * part_total * sizeof(xyzz)) ||
        X.ra_t.ensure((size_t)p.Wb * p.m1 * sizeof(xyzz)) || X.ra_e.ensure((size_t)p.Wb * p.m1 * sizeof(xyzz)) ||
        X.r0.ensure((size_t)p.Wb * p.nb0 * H2_R0_ROWS * sizeof(xyzz)) || X.r1.ensure((size_t)p.Wb * p.r1_rows * sizeof(xyzz)) ||
        X.wsum.ensure((size_t)p.Wb * sizeof(xyzz)))
        return 1;
    MsmBuffers Mk[H2_MAX_UPLOAD_CHUNKS];
    for (uint32_t j = 0; j < K; j++) {
        MsmBuffers &M = Mk[j];
        const size_t o = first[j];
        M.scalars = d_scalars + o; M.bases = d_bases + o; M.bases_phi = X.bases_phi.as<affine>() + o;
        M.glv_parts = X.glv_parts.as<uint32_t>() + 8 * o; M.scalars_mont = scalars_mont ? 1u : 0u;
        M.scal_canon = X.scal_canon.as<fe>() + o;
        M.counts = X.counts.as<uint32_t>() + j * (p.G + 1); M.cursor = X.cursor.as<uint32_t>() + j * 2 * p.G; M.cursor2 = M.cursor + p.G;
        M.refs = X.refs.as<uint32_t>() + j * ref_space;
        M.size_hist = X.size_hist.as<uint32_t>() + j * small_words; M.size_cursor = M.size_hist + (pk[j].T + 2); M.flags = M.size_cursor + (pk[j].T + 2);
        M.items = X.items.as<uint2>() + j * max_items;
        M.bucket_sum = X.bucket_sum.as<xyzz>() + j * p.G;
        M.pkey = X.pkey.as<uint32_t>() + j * part_total; M.pstart = X.pstart.as<uint32_t>() + j * part_total;
        M.pend = X.pend.as<uint32_t>() + j * part_total; M.ppt = X.ppt.as<xyzz>() + j * part_total;
        M.ra_t = X.ra_t.as<xyzz>(); M.ra_e = X.ra_e.as<xyzz>(); M.r0 = X.r0.as<xyzz>(); M.r1 = X.r1.as<xyzz>();
        M.wsum = X.wsum.as<xyzz>(); M.result = d_out;
    }
    X.last_flags = Mk[0].flags;

    {   // scratch of the scan (sized here so that nothing allocates while a graph is being captured)
        const uint64_t per_block = (uint64_t)H2_SCAN_BLOCK * H2_SCAN_ITEMS;
        if (X.scan_blocks.ensure((size_t)((p.G + 1 + per_block - 1) / per_block) * 4 + 16)) return 1;
    }
    issue = [&]() -> int {
        CU(cudaMemsetAsync(X.counts.p, 0, K * (p.G + 1) * 4, s));
        CU(cudaMemsetAsync(X.cursor.p, 0, K * 2 * p.G * 4, s));
        CU(cudaMemsetAsync(X.size_hist.p, 0, K * small_words * 4, s));
        CU(cudaMemsetAsync(X.bucket_sum.p, 0, K * p.G * sizeof(xyzz), s));
        CU(cudaMemsetAsync(X.pkey.p, 0xff, K * part_total * 4, s));

        auto k_bin = msm_bin_kernel<P, PS>;
        auto k_hist = msm_hist_kernel<P, PS>;
        auto k_scatter = msm_scatter_kernel<P, PS>;
        auto k_ihist = msm_item_hist_kernel<P, PS>;
        auto k_ibases = msm_item_bases_kernel<P, PS>;
        auto k_iplace = msm_item_place_kernel<P, PS>;
        auto k_accum0 = msm_accum0_kernel<P, PS>;
        auto k_accum0q = msm_accum0_quad_kernel<P, PS>;
        auto k_accum0m2 = msm_accum0_multi_kernel<P, PS, 2>;
        auto k_accum0m4 = msm_accum0_multi_kernel<P, PS, 4>;
        auto k_accumN = msm_accumN_kernel<P, PS>;
        auto k_rest = msm_accum_rest_kernel<P, PS>;
        auto k_reduceA = msm_reduceA_kernel<P, PS>;
        auto k_r0 = msm_r0_kernel<P, PS>;
        auto k_r1 = msm_r1_kernel<P, PS>;
        auto k_wsum = msm_wsum_kernel<P, PS>;
        auto k_final = msm_final_kernel<P, PS>;
        for (uint32_t j = 0; j < K; j++) {
            const MsmPlan &q = pk[j];
            const MsmBuffers &M = Mk[j];
            if (bc && bc->k) {   // the scalars of this chunk (K == 1: of every chunk of the upload)
                for (uint32_t e = (K > 1 ? j : 0); e < (K > 1 ? j + 1 : bc->k); e++) CU(cudaStreamWaitEvent(s, bc->ev_scal[e], 0));
            }
            // K2/K3: the (point, window) references sorted by bucket -- a single pass into per-bucket bins; the exact
            // histogram / scan / scatter kernels run only if a bin overflowed (flags[1], set by the bin kernel) or if there
            // are no bins (set here)
            if (q.cap == 0) CU(cudaMemsetAsync(M.flags + 1, 0x01, 4, s));
            else LAUNCH(k_bin, blocks_for(q.n * q.sets, 256), 256, 0, s, q, M);
            LAUNCH(k_hist, blocks_for(q.n * q.sets, 256), 256, 0, s, q, M);
            if (exclusive_scan_u32(M.counts, q.G + 1, s, M.flags + 1)) return 1;
            LAUNCH(k_scatter, blocks_for(q.n * q.sets, 256), 256, 0, s, q, M);
            // K4: work items (one per bucket, oversized buckets split), largest first
            LAUNCH(k_ihist, blocks_for(q.G, 256), 256, 0, s, q, M);
            LAUNCH(k_ibases, 1, 32, 0, s, q, M);
            LAUNCH(k_iplace, blocks_for(q.G, 256), 256, 0, s, q, M);
            if (bc && bc->k) {   // the sort above only needed the scalars
                for (uint32_t e = (K > 1 ? j : 0); e < (K > 1 ? j + 1 : bc->k); e++) CU(cudaStreamWaitEvent(s, bc->ev[e], 0));
            }
            if (q.glv) {
                auto k_phi = msm_phi_kernel<P, PS>;
                LAUNCH(k_phi, blocks_for(q.n, 256), 256, 0, s, M.bases, M.bases_phi, (uint64_t)q.n);
            }
            prof_begin(PROF_MSM_ACCUM0, s);
            if (q.max_refs <= H2_MSM_QUAD_ACCUM_REFS) {   // latency-bound: quads, several per work item
                if (X.accum_ways == 4) LAUNCH(k_accum0m4, blocks_for(q.max_items * 16, 128), 128, 0, s, q, M);
                else if (X.accum_ways == 2) LAUNCH(k_accum0m2, blocks_for(q.max_items * 8, 128), 128, 0, s, q, M);
                else LAUNCH(k_accum0q, blocks_for(q.max_items * 4, 128), 128, 0, s, q, M);
            }
            else LAUNCH(k_accum0, blocks_for(q.max_items, 128), 128, 0, s, q, M);
            prof_end(s);
            if (q.acc_levels > 1) LAUNCH(k_accumN, blocks_for(q.acc_threads[1], 128), 128, 0, s, q, M, 1u);
            if (q.acc_levels > 2) LAUNCH(k_accumN, blocks_for(q.acc_threads[2], 128), 128, 0, s, q, M, 2u);
            if (q.acc_levels > 3) LAUNCH(k_rest, 1, 256, 0, s, q, M);
        }
        // K5: bucket reduce (adds the per-chunk bucket sums) and window combine
        const MsmBuffers &M = Mk[0];
        LAUNCH(k_reduceA, blocks_for((uint64_t)p.Wb * p.m1 * 4, 128), 128, 0, s, p, M);                    // quads
        LAUNCH(k_r0, blocks_for((uint64_t)p.Wb * p.nb0 * (2 + p.bits0) * 4, 128), 128, 0, s, p, M);
        LAUNCH(k_r1, p.Wb * p.r1_rows, 4 * H2_R1_QUADS, 0, s, p, M);
        LAUNCH(k_wsum, p.Wb, 128, 0, s, p, M);
        LAUNCH(k_final, 1, 64, 0, s, p, M, (uint32_t)out_canonical);
        return 0;
    };
    return msm_issue_or_replay(issue, fixed && !bc, d_scalars, d_bases, d_out, n, stride, c, sets, scalars_mont, out_canonical, s);
}

static int msm_dispatch(int curve, const fe *d_scalars, int scalars_mont, const affine *d_bases, size_t n, uint32_t c,
                        jacobian *d_out, int out_canonical, cudaStream_t s, uint32_t fixed = 0, uint64_t stride = 0,
                        const BasesChunks *bc = nullptr, uint32_t sets = 1) {
    if (curve == H2_CURVE_PALLAS) return msm_run<FpParams, FqParams>(d_scalars, scalars_mont, d_bases, n, c, fixed, stride, d_out, out_canonical, s, bc, sets);
    if (curve == H2_CURVE_VESTA) return msm_run<FqParams, FpParams>(d_scalars, scalars_mont, d_bases, n, c, fixed, stride, d_out, out_canonical, s, bc, sets);
    return fail("unknown curve id");
}
// window size for a precomputed table over n bases: few references per bucket (short serial chains)
// for small n, fewer windows for large n
static uint32_t table_window(size_t n) {
    uint32_t lg = 0;
    while ((1ull << (lg + 1)) <= n) lg++;
    // candidates are the window sizes whose TOP window is well filled (scalars have 254 significant bits:
    // 254 - (W - 1) c = 6, 14, 14, 16, 14 bits for c = 8, 15, 16, 17, 20): a top window of 1-2 bits would send n / 4
    // references to a handful of shared buckets and defeat the single-pass sort
    uint32_t want = lg + 2;
    if (want <= 9) return 8;
    if (want <= 15) return 15;
    if (want == 16) return 15;        // k = 14: 15 measured better than 16 (IPA opening 5.5 vs 6.2 ms, commit equal; tools/table_sweep.py)
    if (want <= 18) return 17;
    return 20;
}
static int build_table(BaseSet *b, uint32_t c, cudaStream_t s) {
    if (c == 0) c = table_window(b->n);
    if (c < 4 || c > 24) return fail("window table: window bits must be in [4, 24]");
    uint32_t W = (256 + c - 1) / c;
    if ((uint64_t)W * b->n >= (1ull << 31)) return fail("window table: too many points");
    if (b->table.ensure((size_t)W * b->n * sizeof(affine))) return 1;
    if (b->curve == H2_CURVE_PALLAS) {
        auto k = msm_table_kernel<FpParams, FqParams>;
        LAUNCH(k, blocks_for(b->n, 128), 128, 0, s, b->buf.as<affine>(), b->table.as<affine>(), (uint64_t)b->n, (uint64_t)b->n, c, W);
    } else {
        auto k = msm_table_kernel<FqParams, FpParams>;
        LAUNCH(k, blocks_for(b->n, 128), 128, 0, s, b->buf.as<affine>(), b->table.as<affine>(), (uint64_t)b->n, (uint64_t)b->n, c, W);
    }
    b->c = c; b->W = W;
    return 0;
}
// digit-multiples table of a small resident set (fixedbase.cuh), from the c = 8 window table
#define H2_FB_MAX_POINTS ((1u << 15) + 2u)
static int build_direct(BaseSet *b, cudaStream_t s) {
    if (b->n > H2_FB_MAX_POINTS) return fail("H2_BASES_DIRECT: at most 2^15 + 2 points (256 KiB of table per point)");
    if (b->c != H2_FB_BITS || b->W != H2_FB_WINDOWS) return fail("H2_BASES_DIRECT: needs the 8-bit window table");
    if (b->dtable.ensure((size_t)H2_FB_WINDOWS * H2_FB_MULTIPLES * b->n * sizeof(affine))) return 1;
    const uint64_t threads = (uint64_t)H2_FB_WINDOWS * b->n;
    if (b->curve == H2_CURVE_PALLAS) {
        auto k = fb_table_kernel<FpParams, FqParams>;
        LAUNCH(k, blocks_for(threads, 128), 128, 0, s, (const affine *)b->table.as<affine>(), b->dtable.as<affine>(), (uint64_t)b->n, (uint64_t)b->n);
    } else {
        auto k = fb_table_kernel<FqParams, FpParams>;
        LAUNCH(k, blocks_for(threads, 128), 128, 0, s, (const affine *)b->table.as<affine>(), b->dtable.as<affine>(), (uint64_t)b->n, (uint64_t)b->n);
    }
    return 0;
}
// what a fixed-base MSM over `b` runs on: the digit-multiples table (mode 2) when there is one, else the window table (mode 1)
static inline const affine *fixed_table(const BaseSet *b, uint32_t *c, uint32_t *mode) {
    if (b->dtable.p) { *c = H2_FB_BITS; *mode = 2; return b->dtable.as<affine>(); }
    *c = b->c; *mode = 1;
    return b->table.as<affine>();
}
static int convert_points(int curve, affine *d, size_t n, int to_mont, cudaStream_t s) {
    if (n == 0) return 0;
    if (curve == H2_CURVE_PALLAS) LAUNCH(convert_points_kernel<FpParams>, blocks_for(n, 256), 256, 0, s, d, (uint64_t)n, to_mont);
    else LAUNCH(convert_points_kernel<FqParams>, blocks_for(n, 256), 256, 0, s, d, (uint64_t)n, to_mont);
    return 0;
}

extern "C" int h2_msm_dev(int curve, const void *d_scalars, int scalars_repr, const void *d_bases, size_t n, uint32_t window_bits,
                          void *d_out_xyz, void *stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    cudaStream_t s = (cudaStream_t)stream;
    if (scratch_acquire(s)) return 1;
    int rc = msm_dispatch(curve, (const fe *)d_scalars, scalars_repr == H2_REPR_MONTGOMERY, (const affine *)d_bases, n, window_bits,
                          (jacobian *)d_out_xyz, 0, s);
    if (rc) return rc;
    return scratch_release(s);
}

// host_bases != nullptr: one-shot MSM -- the bases are uploaded (and converted) on the copy stream AFTER the
// scalars, overlapping the digit/sort kernels, which only read scalars.
static int msm_host_common(int curve, const void *scalars, size_t n_scalars, const void *extra_scalar, const affine *d_bases,
                           size_t n_total, int repr, void *out_xyz, uint32_t c = 0, uint32_t fixed = 0, uint64_t stride = 0,
                           const void *host_bases = nullptr) {
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    if (scratch_acquire(s)) return 1;
    if (X.scal_in.ensure((n_total + 1) * sizeof(fe)) || X.result.ensure(sizeof(jacobian))) return 1;
    BasesChunks bc;
    if (host_bases && n_total) {
        // One-shot MSM: everything goes up on the copy stream, interleaved per chunk -- scalars of chunk j, then its
        // bases -- so that the sort of chunk j starts when its scalars have landed and its accumulation when its bases
        // have, while chunk j + 1 is on the link.
        // (2 chunks from 2^chunk_min_log points, 3 from 2x that, 4 from 8x: every chunk pays its own sort / work-item launches)
        cudaStream_t cs = X.copy_stream;
        CU(cudaEventRecord(X.ev_scalars_up, s));
        CU(cudaStreamWaitEvent(cs, X.ev_scalars_up, 0));      // after the prior users of the scratch buffers
        bc.k = 1;
        if (n_total >= ((size_t)1 << X.chunk_min_log) && n_total >= 16 * H2_MAX_UPLOAD_CHUNKS)
            bc.k = n_total >= ((size_t)8 << X.chunk_min_log) ? H2_MAX_UPLOAD_CHUNKS : n_total >= ((size_t)2 << X.chunk_min_log) ? 3u : 2u;
        affine *db = const_cast<affine *>(d_bases);
        for (uint32_t j = 0; j < bc.k; j++) {
            size_t lo = chunk_first(n_total, bc.k, j), hi = chunk_first(n_total, bc.k, j + 1);
            CU(cudaMemcpyAsync(X.scal_in.as<fe>() + lo, (const fe *)scalars + lo, (hi - lo) * sizeof(fe), cudaMemcpyHostToDevice, cs));
            CU(cudaEventRecord(X.ev_scal_up[j], cs));
            bc.ev_scal[j] = X.ev_scal_up[j];
            CU(cudaMemcpyAsync(db + lo, (const affine *)host_bases + lo, (hi - lo) * sizeof(affine), cudaMemcpyHostToDevice, cs));
            if (repr == H2_REPR_CANONICAL && convert_points(curve, db + lo, hi - lo, 1, cs)) return 1;
            CU(cudaEventRecord(X.ev_bases_up[j], cs));
            bc.ev[j] = X.ev_bases_up[j];
        }
    } else {
        if (n_scalars) CU(cudaMemcpyAsync(X.scal_in.p, scalars, n_scalars * sizeof(fe), cudaMemcpyHostToDevice, s));
        if (extra_scalar) CU(cudaMemcpyAsync(X.scal_in.as<fe>() + n_scalars, extra_scalar, sizeof(fe), cudaMemcpyHostToDevice, s));
    }
    int rc = msm_dispatch(curve, X.scal_in.as<fe>(), repr == H2_REPR_MONTGOMERY, d_bases, n_total, c, X.result.as<jacobian>(),
                          repr == H2_REPR_CANONICAL, s, fixed, stride, bc.k ? &bc : nullptr);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out_xyz, X.result.p, sizeof(jacobian), cudaMemcpyDeviceToHost, s));
    if (scratch_release(s)) return 1;
    CU(cudaStreamSynchronize(s));
    return 0;
}

extern "C" int h2_msm(int curve, const void *scalars, const void *bases_xy, size_t n, int repr, void *out_xyz) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    if (curve != H2_CURVE_PALLAS && curve != H2_CURVE_VESTA) return fail("unknown curve id");
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    if (scratch_acquire(s)) return 1;
    if (X.bases_in.ensure((n + 1) * sizeof(affine))) return 1;
    return msm_host_common(curve, scalars, n, nullptr, X.bases_in.as<affine>(), n, repr, out_xyz, 0, 0, 0, bases_xy);
}

static int bases_register_impl(int curve, const void *bases_xy, size_t n, int repr, uint32_t window_bits, uint32_t flags, uint64_t *handle);
extern "C" int h2_bases_register(int curve, const void *bases_xy, size_t n, int repr, uint64_t *handle) {
    return bases_register_impl(curve, bases_xy, n, repr, 0, 0, handle);
}
extern "C" int h2_bases_register_ex(int curve, const void *bases_xy, size_t n, int repr, uint32_t window_bits, uint32_t flags, uint64_t *handle) {
    return bases_register_impl(curve, bases_xy, n, repr, window_bits, flags, handle);
}
static int bases_register_impl(int curve, const void *bases_xy, size_t n, int repr, uint32_t window_bits, uint32_t flags, uint64_t *handle) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    if (curve != H2_CURVE_PALLAS && curve != H2_CURVE_VESTA) return fail("unknown curve id");
    BaseSet *b = new BaseSet();
    b->curve = curve; b->n = n;
    if (b->buf.ensure((n + 1) * sizeof(affine))) { delete b; return 1; }
    cudaStream_t s = g_ctx.stream;
    if (n) CU(cudaMemcpyAsync(b->buf.p, bases_xy, n * sizeof(affine), cudaMemcpyHostToDevice, s));
    if (repr == H2_REPR_CANONICAL && convert_points(curve, b->buf.as<affine>(), n, 1, s)) return 1;
    const bool direct = (flags & H2_BASES_DIRECT) && (flags & H2_BASES_PRECOMPUTE) && n > 0;
    if ((flags & H2_BASES_PRECOMPUTE) && n > 0 && build_table(b, direct ? H2_FB_BITS : window_bits, s)) return 1;
    if (direct && build_direct(b, s)) return 1;
    CU(cudaStreamSynchronize(s));
    uint64_t h = g_ctx.next_handle++;
    g_ctx.bases[h] = b;
    *handle = h;
    return 0;
}
extern "C" int h2_bases_release(uint64_t handle) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_ctx.bases.find(handle);
    if (it == g_ctx.bases.end()) return fail("h2_bases_release: unknown handle");
    cudaSetDevice(g_ctx.device);
    cudaDeviceSynchronize();
    it->second->buf.release();
    it->second->table.release();
    it->second->dtable.release();
    delete it->second;
    g_ctx.bases.erase(it);
    return 0;
}
extern "C" int h2_msm_registered(uint64_t handle, const void *scalars, size_t n, const void *extra_scalar, int repr, void *out_xyz) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    auto it = g_ctx.bases.find(handle);
    if (it == g_ctx.bases.end()) return fail("h2_msm_registered: unknown handle");
    BaseSet *b = it->second;
    size_t total = n + (extra_scalar ? 1 : 0);
    if (total > b->n) return fail("h2_msm_registered: more scalars than registered bases");
    if (b->table.p) {   // fixed-base path: digit-multiples table (direct sum) or window table (one shared bucket set)
        uint32_t c, mode;
        const affine *t = fixed_table(b, &c, &mode);
        return msm_host_common(b->curve, scalars, n, extra_scalar, t, total, repr, out_xyz, c, mode, b->n);
    }
    return msm_host_common(b->curve, scalars, n, extra_scalar, b->buf.as<affine>(), total, repr, out_xyz);
}

// `batch` scalar vectors of n entries (+ one extra scalar each, the blinds) against a registered base set with a
// window table: one pass, one bucket set per vector.
static int msm_registered_batch_impl(uint64_t handle, const void *scalars, size_t n, const void *extra_scalars, size_t batch, int repr,
                                     void *out, int affine_out);
extern "C" int h2_msm_registered_batch(uint64_t handle, const void *scalars, size_t n, const void *extra_scalars, size_t batch, int repr,
                                       void *out_xyz) {
    return msm_registered_batch_impl(handle, scalars, n, extra_scalars, batch, repr, out_xyz, 0);
}
// the same pass followed by batch_normalize on the device (plonk/prover.rs:305-311: commit every column, then
// C::Curve::batch_normalize): `batch` affine points (64 B) come back instead of Jacobian ones
extern "C" int h2_msm_registered_batch_affine(uint64_t handle, const void *scalars, size_t n, const void *extra_scalars, size_t batch, int repr,
                                              void *out_xy) {
    return msm_registered_batch_impl(handle, scalars, n, extra_scalars, batch, repr, out_xy, 1);
}
static int msm_registered_batch_impl(uint64_t handle, const void *scalars, size_t n, const void *extra_scalars, size_t batch, int repr,
                                     void *out_xyz, int affine_out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    auto it = g_ctx.bases.find(handle);
    if (it == g_ctx.bases.end()) return fail("h2_msm_registered_batch: unknown handle");
    BaseSet *b = it->second;
    if (!b->table.p) return fail("h2_msm_registered_batch: the base set has no window table (register with H2_BASES_PRECOMPUTE)");
    if (batch == 0) return 0;
    if (batch > 64) return fail("h2_msm_registered_batch: batch > 64");
    size_t total = n + (extra_scalars ? 1 : 0);
    if (total > b->n) return fail("h2_msm_registered_batch: more scalars than registered bases");
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    if (scratch_acquire(s)) return 1;
    if (X.scal_in.ensure(batch * total * sizeof(fe)) || X.result.ensure(batch * sizeof(jacobian))) return 1;
    fe *d = X.scal_in.as<fe>();
    if (!extra_scalars) {
        CU(cudaMemcpyAsync(d, scalars, batch * n * sizeof(fe), cudaMemcpyHostToDevice, s));
    } else {   // interleave: [poly_k (n) | blind_k] per vector
        CU(cudaMemcpy2DAsync(d, total * sizeof(fe), scalars, n * sizeof(fe), n * sizeof(fe), batch, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpy2DAsync(d + n, total * sizeof(fe), extra_scalars, sizeof(fe), sizeof(fe), batch, cudaMemcpyHostToDevice, s));
    }
    const int canon = repr == H2_REPR_CANONICAL;
    uint32_t tc, tmode;
    const affine *tbl = fixed_table(b, &tc, &tmode);
    int rc = msm_dispatch(b->curve, d, repr == H2_REPR_MONTGOMERY, tbl, total, tc, X.result.as<jacobian>(),
                          affine_out ? 0 : canon, s, tmode, b->n, nullptr, (uint32_t)batch);
    if (rc) return rc;
    if (affine_out) {
        if (X.ec_out.ensure(batch * sizeof(affine))) return 1;
        const uint32_t nb = blocks_for((batch + H2_NORM_CHUNK - 1) / H2_NORM_CHUNK, 64);
        if (b->curve == H2_CURVE_PALLAS)
            LAUNCH(normalize_kernel<FpParams>, nb, 64, 0, s, (const xyzz *)nullptr, X.result.as<jacobian>(), 0, X.ec_out.as<affine>(), canon, (uint64_t)batch);
        else
            LAUNCH(normalize_kernel<FqParams>, nb, 64, 0, s, (const xyzz *)nullptr, X.result.as<jacobian>(), 0, X.ec_out.as<affine>(), canon, (uint64_t)batch);
        CU(cudaMemcpyAsync(out_xyz, X.ec_out.p, batch * sizeof(affine), cudaMemcpyDeviceToHost, s));
    } else {
        CU(cudaMemcpyAsync(out_xyz, X.result.p, batch * sizeof(jacobian), cudaMemcpyDeviceToHost, s));
    }
    if (scratch_release(s)) return 1;
    CU(cudaStreamSynchronize(s));
    return 0;
}

extern "C" int h2_point_sum(int curve, const void *points_xyz, size_t g, int repr, void *out_xyz) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    if (scratch_acquire(s)) return 1;
    if (X.misc.ensure((g + 1) * sizeof(jacobian)) || X.result.ensure(sizeof(jacobian))) return 1;
    if (g) CU(cudaMemcpyAsync(X.misc.p, points_xyz, g * sizeof(jacobian), cudaMemcpyHostToDevice, s));
    int canon = repr == H2_REPR_CANONICAL;
    if (curve == H2_CURVE_PALLAS) LAUNCH(point_sum_kernel<FpParams>, 1, 32, 0, s, X.misc.as<jacobian>(), (uint32_t)g, canon, X.result.as<jacobian>());
    else if (curve == H2_CURVE_VESTA) LAUNCH(point_sum_kernel<FqParams>, 1, 32, 0, s, X.misc.as<jacobian>(), (uint32_t)g, canon, X.result.as<jacobian>());
    else return fail("unknown curve id");
    CU(cudaMemcpyAsync(out_xyz, X.result.p, sizeof(jacobian), cudaMemcpyDeviceToHost, s));
    if (scratch_release(s)) return 1;
    CU(cudaStreamSynchronize(s));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// NTT pipeline
// ------------------------------------------------------------------------------------------------
template <class P> static fe host_to_mont(const void *bytes, int repr) {
    fe x;
    memcpy(x.v, bytes, 32);
    return repr == H2_REPR_MONTGOMERY ? x : fe_to_mont<P>(x);
}

template <class P> static int get_twiddles(int field, const fe &omega_mont, uint32_t log_n, cudaStream_t s, const fe **out) {
    Context &X = g_ctx;
    fe canon = fe_from_mont<P>(omega_mont);
    for (auto *t : X.twiddles)
        if (t->field == field && t->log_n == log_n && memcmp(t->omega, canon.v, 32) == 0) {
            t->stamp = ++X.tw_stamp;
            *out = t->buf.as<fe>();
            return 0;
        }
    TwiddleEntry *e = nullptr;
    if (X.twiddles.size() >= 8) {   // evict least recently used
        size_t victim = 0;
        for (size_t i = 1; i < X.twiddles.size(); i++)
            if (X.twiddles[i]->stamp < X.twiddles[victim]->stamp) victim = i;
        e = X.twiddles[victim];
        X.twiddles.erase(X.twiddles.begin() + victim);
        CU(cudaStreamSynchronize(s));
    } else e = new TwiddleEntry();
    uint64_t half = log_n ? (1ull << (log_n - 1)) : 1;
    if (e->buf.ensure(half * sizeof(fe)) || X.pow2.ensure(64 * sizeof(fe))) { delete e; return 1; }
    e->field = field; e->log_n = log_n; memcpy(e->omega, canon.v, 32); e->stamp = ++X.tw_stamp;
    LAUNCH(twiddle_pow2_kernel<P>, 1, 32, 0, s, X.pow2.as<fe>(), omega_mont, log_n ? log_n : 1u);
    LAUNCH(twiddle_fill_kernel<P>, blocks_for((half + 31) / 32, 128), 128, 0, s, e->buf.as<fe>(), X.pow2.as<fe>(), half);
    X.twiddles.push_back(e);
    *out = e->buf.as<fe>();
    return 0;
}

struct NttScales {
    bool in_scale = false, out_scale = false;
    fe in_s[3], out_s[3];
};

// d_in: 2^in_log_n elements; d_out: min(out_len, 2^log_n) elements written.  d_out may alias d_in.
template <class P>
static int ntt_run(int field, const fe *d_in, uint32_t in_log_n, fe *d_out, uint32_t log_n, const fe &omega_mont, const NttScales &sc,
                   uint64_t out_len, cudaStream_t s) {
    Context &X = g_ctx;
    if (log_n > 30) return fail("ntt: log_n > 30 not supported");
    uint64_t n = 1ull << log_n;
    const fe *tw = nullptr;
    if (get_twiddles<P>(field, omega_mont, log_n, s, &tw)) return 1;
    uint32_t sp[8], logc[8];
    int passes = ntt_plan(log_n, sp, logc);
    if (passes == 0) {   // n == 1: the network is empty; only the scalings apply
        sp[0] = 0; logc[0] = 0; passes = 1;
    }
    if (passes > 1 && X.ntt_work.ensure(n * sizeof(fe))) return 1;
    uint32_t s0 = 0;
    for (int i = 0; i < passes; i++) {
        NttPassArgs A;
        A.in = i == 0 ? d_in : X.ntt_work.as<fe>();
        A.out = i == passes - 1 ? d_out : X.ntt_work.as<fe>();
        A.tw = tw; A.log_n = log_n; A.s0 = s0; A.sp = sp[i]; A.logc = logc[i];
        A.flags = 0;
        if (i == 0) A.flags |= NTT_FIRST | (sc.in_scale ? NTT_IN_SCALE : 0u);
        if (i == passes - 1) A.flags |= NTT_LAST | (sc.out_scale ? NTT_OUT_SCALE : 0u);
        A.in_log_n = in_log_n; A.out_len = out_len;
        for (int k = 0; k < 3; k++) { A.in_scale[k] = sc.in_s[k]; A.out_scale[k] = sc.out_s[k]; }
        uint32_t tiles = (uint32_t)(n >> (sp[i] + logc[i]));
        uint32_t smem = ntt_smem_bytes(sp[i], logc[i]);
        prof_begin(PROF_NTT_PASS, s);
        LAUNCH(ntt_pass_kernel<P>, tiles, 128, smem, s, A);
        prof_end(s);
        s0 += sp[i];
    }
    return 0;
}

// Builds the in/out scale constants.  data_repr is the encoding of the data entering and leaving.
//   zeta_in  != null : multiply element j by zeta^(j mod 3)            (coeff_to_extended)
//   divisor  != null : multiply every output by divisor                (ifft)
//   zeta_out != null : multiply output p by [1, zeta^2, zeta][p mod 3] (extended_to_coeff)
template <class P>
static NttScales make_scales(int data_repr, const fe *zeta_in, const fe *divisor, const fe *zeta_out) {
    NttScales sc;
    fe one = fe_one<P>();
    fe in_c[3] = {one, one, one}, out_c[3] = {one, one, one};
    bool in_needed = false, out_needed = false;
    if (zeta_in) { in_c[1] = *zeta_in; in_c[2] = fe_sqr<P>(*zeta_in); in_needed = true; }
    if (divisor) { for (int k = 0; k < 3; k++) out_c[k] = *divisor; out_needed = true; }
    if (zeta_out) { out_c[1] = fe_mul<P>(out_c[1], fe_sqr<P>(*zeta_out)); out_c[2] = fe_mul<P>(out_c[2], *zeta_out); out_needed = true; }
    if (data_repr == H2_REPR_CANONICAL) {
        // canonical -> Montgomery on the way in:  mont_mul(a, c R^2) = a c R
        for (int k = 0; k < 3; k++) in_c[k] = fe_mul<P>(in_c[k], fe_r2<P>());
        // Montgomery -> canonical on the way out: mont_mul(x R, c) = x c
        for (int k = 0; k < 3; k++) out_c[k] = fe_from_mont<P>(out_c[k]);
        in_needed = out_needed = true;
    }
    sc.in_scale = in_needed; sc.out_scale = out_needed;
    for (int k = 0; k < 3; k++) { sc.in_s[k] = in_c[k]; sc.out_s[k] = out_c[k]; }
    return sc;
}

template <class P>
static int ntt_host(int field, int mode, const void *a_in, uint32_t in_log_n, uint32_t log_n, const void *omega, const void *zeta,
                    const void *divisor, size_t out_len, void *out, int repr) {
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    if (scratch_acquire(s)) return 1;
    uint64_t n = 1ull << log_n, n_in = 1ull << in_log_n;
    if (out_len > n) out_len = n;
    if (X.ntt_io.ensure(n_in * sizeof(fe)) || X.ntt_out.ensure(n * sizeof(fe))) return 1;
    fe w = host_to_mont<P>(omega, repr), z, d;
    if (zeta) z = host_to_mont<P>(zeta, repr);
    if (divisor) d = host_to_mont<P>(divisor, repr);
    NttScales sc = make_scales<P>(repr, mode == 2 ? &z : nullptr, (mode == 1 || mode == 3) ? &d : nullptr, mode == 3 ? &z : nullptr);
    CU(cudaMemcpyAsync(X.ntt_io.p, a_in, n_in * sizeof(fe), cudaMemcpyHostToDevice, s));
    if (ntt_run<P>(field, X.ntt_io.as<fe>(), in_log_n, X.ntt_out.as<fe>(), log_n, w, sc, out_len, s)) return 1;
    CU(cudaMemcpyAsync(out, X.ntt_out.p, out_len * sizeof(fe), cudaMemcpyDeviceToHost, s));
    if (scratch_release(s)) return 1;
    CU(cudaStreamSynchronize(s));
    return 0;
}
static int ntt_host_dispatch(int field, int mode, const void *a_in, uint32_t in_log_n, uint32_t log_n, const void *omega, const void *zeta,
                             const void *divisor, size_t out_len, void *out, int repr) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    if (log_n > 30 || in_log_n > log_n) return fail("ntt: bad sizes");
    if (field == H2_FIELD_FP) return ntt_host<FpParams>(field, mode, a_in, in_log_n, log_n, omega, zeta, divisor, out_len, out, repr);
    if (field == H2_FIELD_FQ) return ntt_host<FqParams>(field, mode, a_in, in_log_n, log_n, omega, zeta, divisor, out_len, out, repr);
    return fail("unknown field id");
}
extern "C" int h2_ntt(int field, void *a, const void *omega, uint32_t log_n, int repr) {
    return ntt_host_dispatch(field, 0, a, log_n, log_n, omega, nullptr, nullptr, (size_t)1 << log_n, a, repr);
}
extern "C" int h2_intt_scaled(int field, void *a, const void *omega_inv, const void *divisor, uint32_t log_n, int repr) {
    return ntt_host_dispatch(field, 1, a, log_n, log_n, omega_inv, nullptr, divisor, (size_t)1 << log_n, a, repr);
}
extern "C" int h2_coeff_to_extended(int field, const void *a, uint32_t k, uint32_t ext_k, const void *zeta, const void *ext_omega,
                                    void *out, int repr) {
    return ntt_host_dispatch(field, 2, a, k, ext_k, ext_omega, zeta, nullptr, (size_t)1 << ext_k, out, repr);
}
extern "C" int h2_extended_to_coeff(int field, const void *a, uint32_t ext_k, const void *ext_omega_inv, const void *ext_divisor,
                                    const void *zeta, size_t out_len, void *out, int repr) {
    return ntt_host_dispatch(field, 3, a, ext_k, ext_k, ext_omega_inv, zeta, ext_divisor, out_len, out, repr);
}
extern "C" int h2_ntt_dev(int field, const void *d_in, void *d_out, const void *omega, int omega_repr, uint32_t log_n, void *stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    cudaStream_t s = (cudaStream_t)stream;
    if (scratch_acquire(s)) return 1;
    NttScales sc;   // Montgomery in, Montgomery out, no scaling
    int rc;
    if (field == H2_FIELD_FP) rc = ntt_run<FpParams>(field, (const fe *)d_in, log_n, (fe *)d_out, log_n, host_to_mont<FpParams>(omega, omega_repr), sc, 1ull << log_n, s);
    else if (field == H2_FIELD_FQ) rc = ntt_run<FqParams>(field, (const fe *)d_in, log_n, (fe *)d_out, log_n, host_to_mont<FqParams>(omega, omega_repr), sc, 1ull << log_n, s);
    else return fail("unknown field id");
    if (rc) return rc;
    return scratch_release(s);
}
extern "C" int h2_ntt_clear_cache(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    cudaDeviceSynchronize();
    for (auto *t : g_ctx.twiddles) { t->buf.release(); delete t; }
    g_ctx.twiddles.clear();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// EC-FFT (best_fft with G = curve point) and batch normalisation (ecfft.cuh)
// ------------------------------------------------------------------------------------------------
// the log n butterfly stages (+ the optional `*g *= scale` pass) on an XYZZ work array already in network order
template <class P, class PS>
static int ecfft_stages(int scalar_field, xyzz *work, uint32_t log_n, const fe &omega_mont, const fe *scale_canon, cudaStream_t s) {
    const fe *tw = nullptr;
    if (get_twiddles<PS>(scalar_field, omega_mont, log_n, s, &tw)) return 1;
    const uint64_t n = 1ull << log_n;
    // one QUAD of lanes per butterfly (ecfft.cuh) unless the test hook asks for the one-thread form
    // Measured (k = 10 / 12 / 14, g -> g_lagrange): quads 10.3 / 12.4 / 18.2 ms, one thread per butterfly 12.6 / 14.8 /
    // 17.6 ms -- a quad level costs ~3 multiply latencies (selects, call, 32 shuffles, limb carries), so the quad form only
    // wins while a stage has too few butterflies to give every SM a warp.  ecfft_quad: 1 = by size (default), 0 / 2 = force
    // the thread / quad form (test hook).
    const bool use_quad = g_ctx.ecfft_quad == 2 || (g_ctx.ecfft_quad == 1 && log_n <= 12);
    const uint32_t q = use_quad ? 4u : 1u;
    auto stage = use_quad ? ecfft_stage_quad_kernel<P, PS> : ecfft_stage_kernel<P, PS>;
    for (uint32_t st = 1; st <= log_n; st++) LAUNCH(stage, blocks_for(n / 2 * q, 64), 64, 0, s, work, tw, log_n, st);
    if (scale_canon) {
        auto sc = use_quad ? ecfft_scale_quad_kernel<P, PS> : ecfft_scale_kernel<P, PS>;
        LAUNCH(sc, blocks_for(n * q, 64), 64, 0, s, work, *scale_canon, n);
    }
    return 0;
}
// mode 0: Jacobian in -> Jacobian out (h2_ec_fft); mode 1: affine in -> scaled, normalised affine out (h2_params_lagrange)
// `in` == nullptr: the input is already in X.ec_io on the device (h2_params_new), scratch acquired by the caller
template <class P, class PS>
static int ecfft_host(int scalar_field, int mode, const void *in, uint32_t log_n, const void *omega, const void *scale, int repr, void *out) {
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    const uint64_t n = 1ull << log_n;
    const int canon = repr == H2_REPR_CANONICAL;
    const size_t in_sz = mode == 0 ? sizeof(jacobian) : sizeof(affine);
    if (in) {
        if (scratch_acquire(s)) return 1;
        if (X.ec_io.ensure(n * sizeof(jacobian))) return 1;
        CU(cudaMemcpyAsync(X.ec_io.p, in, n * in_sz, cudaMemcpyHostToDevice, s));
    }
    if (X.ec_work.ensure(n * sizeof(xyzz)) || X.ec_out.ensure(n * sizeof(affine))) return 1;
    xyzz *work = X.ec_work.as<xyzz>();
    if (mode == 0) {
        auto k = ecfft_load_jac_kernel<P, PS>;
        LAUNCH(k, blocks_for(n, 128), 128, 0, s, X.ec_io.as<jacobian>(), canon, work, log_n);
    } else {
        auto k = ecfft_load_affine_kernel<P, PS>;
        LAUNCH(k, blocks_for(n, 128), 128, 0, s, X.ec_io.as<affine>(), canon, work, log_n);
    }
    fe sc, *scp = nullptr;
    if (scale) {
        memcpy(sc.v, scale, 32);
        if (!canon) sc = fe_from_mont<PS>(sc);
        scp = &sc;
    }
    if (ecfft_stages<P, PS>(scalar_field, work, log_n, host_to_mont<PS>(omega, repr), scp, s)) return 1;
    if (mode == 0) {
        auto k = ecfft_store_jac_kernel<P, PS>;
        LAUNCH(k, blocks_for(n, 128), 128, 0, s, work, X.ec_io.as<jacobian>(), canon, n);
        CU(cudaMemcpyAsync(out, X.ec_io.p, n * sizeof(jacobian), cudaMemcpyDeviceToHost, s));
    } else {
        LAUNCH(normalize_kernel<P>, blocks_for((n + H2_NORM_CHUNK - 1) / H2_NORM_CHUNK, 64), 64, 0, s, work, (const jacobian *)nullptr, 0,
               X.ec_out.as<affine>(), canon, n);
        CU(cudaMemcpyAsync(out, X.ec_out.p, n * sizeof(affine), cudaMemcpyDeviceToHost, s));
    }
    if (scratch_release(s)) return 1;
    CU(cudaStreamSynchronize(s));
    return 0;
}
static int ecfft_host_dispatch(int curve, int mode, const void *in, uint32_t log_n, const void *omega, const void *scale, int repr, void *out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    if (log_n > 26) return fail("ec_fft: log_n > 26 not supported");
    if (curve == H2_CURVE_PALLAS) return ecfft_host<FpParams, FqParams>(H2_FIELD_FQ, mode, in, log_n, omega, scale, repr, out);
    if (curve == H2_CURVE_VESTA) return ecfft_host<FqParams, FpParams>(H2_FIELD_FP, mode, in, log_n, omega, scale, repr, out);
    return fail("unknown curve id");
}
extern "C" int h2_ec_fft(int curve, void *points_xyz, const void *omega, uint32_t log_n, const void *scale, int repr) {
    return ecfft_host_dispatch(curve, 0, points_xyz, log_n, omega, scale, repr, points_xyz);
}
extern "C" int h2_params_lagrange(int curve, const void *g_xy, uint32_t k, const void *omega_inv, const void *minv, int repr, void *out_xy) {
    if (!minv) return fail("h2_params_lagrange: minv is required (poly/commitment.rs:83)");
    return ecfft_host_dispatch(curve, 1, g_xy, k, omega_inv, minv, repr, out_xy);
}

// ------------------------------------------------------------------------------------------------
// hash_to_curve (h2c.cuh) and Params::new (poly/commitment.rs:38-114)
// ------------------------------------------------------------------------------------------------
// n messages -> n affine points in X.ec_io (device, `repr`); scratch held by the caller.  msgs == nullptr: generator
// messages 0 || (first + i) as u32 LE
template <class P>
static int h2c_issue(const H2cConst &K, const void *msgs, size_t msg_len, uint64_t first, size_t n, int repr, affine *d_out, cudaStream_t s) {
    Context &X = g_ctx;
    const uint8_t *d_msgs = nullptr;
    if (msgs && n * msg_len) {
        if (X.misc.ensure(n * msg_len)) return 1;
        CU(cudaMemcpyAsync(X.misc.p, msgs, n * msg_len, cudaMemcpyHostToDevice, s));
        d_msgs = X.misc.as<uint8_t>();
    }
    LAUNCH(h2c_kernel<P>, blocks_for(n, 64), 64, 0, s, d_msgs, (uint32_t)msg_len, msgs ? 0 : 1, first, K, d_out, repr == H2_REPR_MONTGOMERY ? 1 : 0,
           (uint64_t)n);
    return 0;
}
template <class P>
static int hash_to_curve_host(const char *domain_prefix, const void *msgs, size_t msg_len, size_t n, int repr, void *out_xy) {
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    const H2cConst K = make_h2c_const<P>(domain_prefix);
    if (!K.ok) return fail("h2_hash_to_curve: domain prefix too long (DST must be < 256 bytes)");
    if (scratch_acquire(s)) return 1;
    if (X.ec_io.ensure(n * sizeof(affine))) return 1;
    if (h2c_issue<P>(K, msgs, msg_len, 0, n, repr, X.ec_io.as<affine>(), s)) return 1;
    CU(cudaMemcpyAsync(out_xy, X.ec_io.p, n * sizeof(affine), cudaMemcpyDeviceToHost, s));
    if (scratch_release(s)) return 1;
    CU(cudaStreamSynchronize(s));
    return 0;
}
extern "C" int h2_hash_to_curve(int curve, const char *domain_prefix, const void *messages, size_t msg_len, size_t n, int repr, void *out_xy) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    if (curve != H2_CURVE_PALLAS && curve != H2_CURVE_VESTA) return fail("unknown curve id");
    if (!domain_prefix) return fail("h2_hash_to_curve: domain_prefix is NULL");
    if (msg_len && !messages && n) return fail("h2_hash_to_curve: messages is NULL");
    if (msg_len >= (1ull << 31) || n >= (1ull << 32)) return fail("h2_hash_to_curve: message or batch too large");
    if (n == 0) return 0;
    static const uint8_t empty = 0;
    const void *m = messages ? messages : &empty;      // msg_len == 0: n hashes of the empty message
    return curve == H2_CURVE_PALLAS ? hash_to_curve_host<FpParams>(domain_prefix, m, msg_len, n, repr, out_xy)
                                    : hash_to_curve_host<FqParams>(domain_prefix, m, msg_len, n, repr, out_xy);
}
// Params::new: g[i] = H(0 || i), w = H(1), u = H(2) with H = hash_to_curve("Halo2-Parameters"), then g_lagrange from g
template <class P, class PS>
static int params_new_host(int scalar_field, uint32_t k, int repr, void *g_xy, void *gl_xy, void *w_xy, void *u_xy) {
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    const uint64_t n = 1ull << k;
    static const H2cConst K = make_h2c_const<P>("Halo2-Parameters");
    if (!K.ok) return fail("h2_params_new: internal constant check failed");
    // alpha_inv = ROOT_OF_UNITY_INV^(2^(S-k)) (commitment.rs:77-80), minv = TWO_INV^k (:83)
    static const SqrtConst KS = make_sqrt_const<PS>();
    fe alpha_inv = fe_inv<PS>(KS.root);
    for (uint32_t i = k; i < 32; i++) alpha_inv = fe_sqr<PS>(alpha_inv);
    const fe two_inv = fe_inv<PS>(fe_dbl<PS>(fe_one<PS>()));
    fe minv = fe_one<PS>();
    for (uint32_t i = 0; i < k; i++) minv = fe_mul<PS>(minv, two_inv);
    if (repr == H2_REPR_CANONICAL) { alpha_inv = fe_from_mont<PS>(alpha_inv); minv = fe_from_mont<PS>(minv); }
    if (scratch_acquire(s)) return 1;
    if (X.ec_io.ensure((n + 2) * sizeof(jacobian))) return 1;
    affine *d_g = X.ec_io.as<affine>();
    static const uint8_t wu[2] = {1, 2};
    if (h2c_issue<P>(K, wu, 1, 0, 2, repr, d_g + n, s)) return 1;       // w, u behind g
    if (h2c_issue<P>(K, nullptr, 0, 0, n, repr, d_g, s)) return 1;
    CU(cudaMemcpyAsync(g_xy, d_g, n * sizeof(affine), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(w_xy, d_g + n, sizeof(affine), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(u_xy, d_g + n + 1, sizeof(affine), cudaMemcpyDeviceToHost, s));
    return ecfft_host<P, PS>(scalar_field, 1, nullptr, k, alpha_inv.v, minv.v, repr, gl_xy);   // releases the scratch, synchronises
}
extern "C" int h2_params_new(int curve, uint32_t k, int repr, void *out_g_xy, void *out_g_lagrange_xy, void *out_w_xy, void *out_u_xy) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    if (k > 26) return fail("h2_params_new: k > 26 not supported");
    if (!out_g_xy || !out_g_lagrange_xy || !out_w_xy || !out_u_xy) return fail("h2_params_new: NULL output");
    if (curve == H2_CURVE_PALLAS) return params_new_host<FpParams, FqParams>(H2_FIELD_FQ, k, repr, out_g_xy, out_g_lagrange_xy, out_w_xy, out_u_xy);
    if (curve == H2_CURVE_VESTA) return params_new_host<FqParams, FpParams>(H2_FIELD_FP, k, repr, out_g_xy, out_g_lagrange_xy, out_w_xy, out_u_xy);
    return fail("unknown curve id");
}
extern "C" int h2_batch_normalize(int curve, const void *points_xyz, size_t n, int repr, void *out_xy) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    if (curve != H2_CURVE_PALLAS && curve != H2_CURVE_VESTA) return fail("unknown curve id");
    if (n == 0) return 0;
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    const int canon = repr == H2_REPR_CANONICAL;
    if (scratch_acquire(s)) return 1;
    if (X.ec_io.ensure(n * sizeof(jacobian)) || X.ec_out.ensure(n * sizeof(affine))) return 1;
    CU(cudaMemcpyAsync(X.ec_io.p, points_xyz, n * sizeof(jacobian), cudaMemcpyHostToDevice, s));
    const uint32_t nb = blocks_for((n + H2_NORM_CHUNK - 1) / H2_NORM_CHUNK, 64);
    if (curve == H2_CURVE_PALLAS)
        LAUNCH(normalize_kernel<FpParams>, nb, 64, 0, s, (const xyzz *)nullptr, X.ec_io.as<jacobian>(), canon, X.ec_out.as<affine>(), canon, (uint64_t)n);
    else
        LAUNCH(normalize_kernel<FqParams>, nb, 64, 0, s, (const xyzz *)nullptr, X.ec_io.as<jacobian>(), canon, X.ec_out.as<affine>(), canon, (uint64_t)n);
    CU(cudaMemcpyAsync(out_xy, X.ec_out.p, n * sizeof(affine), cudaMemcpyDeviceToHost, s));
    if (scratch_release(s)) return 1;
    CU(cudaStreamSynchronize(s));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// point (de)compression (codec.cuh): C::to_bytes / C::from_bytes, the encoding of Params::{write, read} and of every
// point in a proof transcript
// ------------------------------------------------------------------------------------------------
template <class P> static int points_codec(int decompress, const void *in, size_t n, int repr, void *out) {
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    const int mont = repr == H2_REPR_MONTGOMERY;
    if (scratch_acquire(s)) return 1;
    if (X.ec_io.ensure(n * sizeof(affine)) || X.ec_out.ensure(n * sizeof(affine)) || X.misc.ensure(64)) return 1;
    uint32_t bad = 0xffffffffu;
    if (!decompress) {
        CU(cudaMemcpyAsync(X.ec_io.p, in, n * sizeof(affine), cudaMemcpyHostToDevice, s));
        LAUNCH(compress_kernel<P>, blocks_for(n, 128), 128, 0, s, X.ec_io.as<affine>(), mont, X.ec_out.as<fe>(), (uint64_t)n);
        CU(cudaMemcpyAsync(out, X.ec_out.p, n * sizeof(fe), cudaMemcpyDeviceToHost, s));
    } else {
        static const SqrtConst K = make_sqrt_const<P>();
        CU(cudaMemcpyAsync(X.ec_io.p, in, n * sizeof(fe), cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(X.misc.p, &bad, 4, cudaMemcpyHostToDevice, s));
        LAUNCH(decompress_kernel<P>, blocks_for(n, 128), 128, 0, s, X.ec_io.as<fe>(), X.ec_out.as<affine>(), mont, K, X.misc.as<uint32_t>(), (uint64_t)n);
        CU(cudaMemcpyAsync(out, X.ec_out.p, n * sizeof(affine), cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(&bad, X.misc.p, 4, cudaMemcpyDeviceToHost, s));
    }
    if (scratch_release(s)) return 1;
    CU(cudaStreamSynchronize(s));
    if (bad != 0xffffffffu) return fail("h2_points_decompress: invalid point encoding at index " + std::to_string(bad));
    return 0;
}
static int points_codec_dispatch(int curve, int decompress, const void *in, size_t n, int repr, void *out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    if (curve != H2_CURVE_PALLAS && curve != H2_CURVE_VESTA) return fail("unknown curve id");
    if (n >= (1ull << 32)) return fail("points codec: n >= 2^32");
    if (n == 0) return 0;
    return curve == H2_CURVE_PALLAS ? points_codec<FpParams>(decompress, in, n, repr, out) : points_codec<FqParams>(decompress, in, n, repr, out);
}
extern "C" int h2_points_compress(int curve, const void *points_xy, size_t n, int repr, void *out_bytes) {
    return points_codec_dispatch(curve, 0, points_xy, n, repr, out_bytes);
}
extern "C" int h2_points_decompress(int curve, const void *bytes, size_t n, int repr, void *out_xy) {
    return points_codec_dispatch(curve, 1, bytes, n, repr, out_xy);
}

// ------------------------------------------------------------------------------------------------
// utilities
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// IPA round loop (ipa.cuh): poly/commitment/prover.rs:100-142 with resident generators
// ------------------------------------------------------------------------------------------------
// eval_polynomial / compute_inner_product / kate_division on resident polynomials (polyops.cuh)
// ------------------------------------------------------------------------------------------------
// mode 0: eval (points: batch x 32 host), 1: inner product of a[i] and c[i], 2: kate division of a[i] by (X - point_i) into c[i]
static PolyBuf *find_poly(uint64_t h);
static int convert_field(int field, fe *d, size_t n, int to_mont, cudaStream_t s);
template <class P>
static int polyops_run(int mode, const std::vector<PolyBuf *> &a, const std::vector<PolyBuf *> &c, size_t n, const void *points, int repr, void *out) {
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    const uint32_t batch = (uint32_t)a.size();
    // level sizes: m[0] = n, m[l + 1] = ceil(m[l] / CHUNK), down to 1
    std::vector<uint64_t> m{(uint64_t)n}, off{0};
    while (m.back() > 1) { off.push_back(off.back() + (m.size() > 1 ? m.back() * batch : 0)); m.push_back((m.back() + H2_POLY_CHUNK - 1) / H2_POLY_CHUNK); }
    if (mode == 1 && m.size() == 1) { off.push_back(0); m.push_back(1); }   // a length-1 inner product still needs its product level
    const size_t L = m.size() - 1;                               // levels above the polynomial itself
    const uint64_t lvl_total = off.back() + m.back() * batch + batch;
    if (scratch_acquire(s)) return 1;
    if (X.po_lvl.ensure(lvl_total * sizeof(fe)) || X.po_q.ensure(lvl_total * sizeof(fe)) || X.po_pts.ensure((L + 2) * batch * sizeof(fe)) ||
        X.po_ptrs.ensure(2 * batch * sizeof(void *)) || X.misc.ensure(batch * sizeof(fe) + 64))
        return 1;
    fe *lvl = X.po_lvl.as<fe>(), *qarr = X.po_q.as<fe>(), *pts = X.po_pts.as<fe>();
    auto level = [&](size_t l) { return lvl + off[l]; };         // values of level l >= 1: [batch][m[l]]
    auto qlevel = [&](size_t l) { return qarr + off[l]; };
    std::vector<const fe *> hp(2 * batch);
    for (uint32_t b = 0; b < batch; b++) { hp[b] = a[b]->buf.as<fe>(); hp[batch + b] = c.empty() ? nullptr : c[b]->buf.as<fe>(); }
    CU(cudaMemcpyAsync(X.po_ptrs.p, hp.data(), 2 * batch * sizeof(void *), cudaMemcpyHostToDevice, s));
    const fe *const *d_a = X.po_ptrs.as<const fe *>();
    const fe *const *d_c = d_a + batch;
    // points of level 0 (Montgomery): the caller's, or 1 for the plain sums of the inner product
    if (mode == 1) {
        LAUNCH(fe_fill_kernel<P>, blocks_for(batch, 64), 64, 0, s, pts, batch, fe_one<P>());
    } else {
        CU(cudaMemcpyAsync(pts, points, batch * sizeof(fe), cudaMemcpyHostToDevice, s));
        if (repr == H2_REPR_CANONICAL) LAUNCH(convert_kernel<P>, blocks_for(batch, 64), 64, 0, s, pts, (uint64_t)batch, 1);
    }
    // upward pass: level l + 1 from level l at the point x^(CHUNK^l)
    for (size_t l = 0; l < L; l++) {
        const dim3 grid(blocks_for(m[l + 1], 128), batch);
        if (l == 0 && mode == 1) LAUNCH(poly_inner_level0_kernel<P>, grid, 128, 0, s, d_a, d_c, m[0], level(1), m[1]);
        else LAUNCH(poly_eval_level_kernel<P>, grid, 128, 0, s, l == 0 ? d_a : (const fe *const *)nullptr, l == 0 ? (const fe *)nullptr : (const fe *)level(l),
                    m[l], (const fe *)(pts + l * batch), level(l + 1), m[l + 1]);
        if (mode != 1) LAUNCH(poly_pow_chunk_kernel<P>, blocks_for(batch, 64), 64, 0, s, (const fe *)(pts + l * batch), pts + (l + 1) * batch, batch);
        else if (l == 0) LAUNCH(fe_fill_kernel<P>, blocks_for(batch, 64), 64, 0, s, pts + batch, batch, fe_one<P>());
        if (mode == 1 && l >= 1) CU(cudaMemcpyAsync(pts + (l + 1) * batch, pts, batch * sizeof(fe), cudaMemcpyDeviceToDevice, s));
    }
    if (mode != 2) {   // the single value of the top level is the result (n == 1: the coefficient itself; n == 0 handled by the caller)
        fe *res = X.misc.as<fe>();
        if (L == 0) {
            for (uint32_t b = 0; b < batch; b++) CU(cudaMemcpyAsync(res + b, hp[b], sizeof(fe), cudaMemcpyDeviceToDevice, s));
        } else {
            CU(cudaMemcpyAsync(res, level(L), batch * sizeof(fe), cudaMemcpyDeviceToDevice, s));   // m[L] == 1: [batch][1]
        }
        if (repr == H2_REPR_CANONICAL) LAUNCH(convert_kernel<P>, blocks_for(batch, 64), 64, 0, s, res, (uint64_t)batch, 0);
        CU(cudaMemcpyAsync(out, res, batch * sizeof(fe), cudaMemcpyDeviceToHost, s));
        if (scratch_release(s)) return 1;
        CU(cudaStreamSynchronize(s));
        return 0;
    }
    // kate division, downward pass: Q at every position of level l from the carries of level l + 1.  The top level with
    // more than one value (m[L] == 1 always; start from the highest level that has something to walk) needs no carry.
    fe *const *d_q = (fe *const *)d_c;
    for (size_t l = L; l-- > 0;) {
        const dim3 grid(blocks_for(m[l + 1], 128), batch);
        const fe *carry = (l + 1 < L) ? (const fe *)qlevel(l + 1) : (const fe *)nullptr;   // Q of level l + 1; the top level's Q(1..) are zero
        LAUNCH(poly_kate_down_kernel<P>, grid, 128, 0, s, l == 0 ? d_a : (const fe *const *)nullptr, l == 0 ? (const fe *)nullptr : (const fe *)level(l), m[l],
               (const fe *)(pts + l * batch), carry, m[l + 1], l == 0 ? (fe *)nullptr : qlevel(l), l == 0 ? d_q : (fe *const *)nullptr);
    }
    return scratch_release(s);
}
static int polyops_dispatch(int mode, const uint64_t *ah, const uint64_t *ch, size_t batch, size_t n, const void *points, int repr, void *out,
                            const char *who) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    if (batch == 0) return 0;
    if (batch > 256) return fail(std::string(who) + ": batch > 256");
    if (n >= (1ull << 32)) return fail(std::string(who) + ": n >= 2^32");
    std::vector<PolyBuf *> a(batch), c;
    if (ch) c.resize(batch);
    for (size_t b = 0; b < batch; b++) {
        a[b] = find_poly(ah[b]);
        if (!a[b]) return fail(std::string(who) + ": unknown polynomial handle");
        if (a[b]->field != a[0]->field) return fail(std::string(who) + ": the polynomials live in different fields");
        if (a[b]->len < n) return fail(std::string(who) + ": a polynomial holds fewer than n coefficients");
        if (ch) {
            c[b] = find_poly(ch[b]);
            if (!c[b]) return fail(std::string(who) + ": unknown polynomial handle");
            if (c[b]->field != a[0]->field) return fail(std::string(who) + ": the polynomials live in different fields");
            if (c[b]->len + (mode == 2 ? 1 : 0) < n) return fail(std::string(who) + ": the second polynomial is too short");
            if (mode == 2 && c[b] == a[b]) return fail(std::string(who) + ": the quotient cannot overwrite its dividend");
        }
    }
    if (a[0]->field == H2_FIELD_FP) return polyops_run<FpParams>(mode, a, c, n, points, repr, out);
    return polyops_run<FqParams>(mode, a, c, n, points, repr, out);
}
// Evaluator::evaluate (poly/evaluator.rs:129-228) on resident polynomials: `code` is the postfix form of the Ast (asteval.cuh),
// validated here so that the kernel's operand stack can neither overflow nor underflow.
template <class P>
static int ast_run(PolyBuf *out, const std::vector<PolyBuf *> &polys, uint32_t log_n, const AstInstr *code, size_t n_code, const void *consts,
                   size_t n_consts, const void *omega, const void *lin_base, int repr, bool has_linear) {
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    const uint64_t n = 1ull << log_n;
    if (scratch_acquire(s)) return 1;
    if (X.ast_code.ensure(n_code * sizeof(AstInstr)) || X.ast_consts.ensure((n_consts + 1) * sizeof(fe)) || X.po_ptrs.ensure((polys.size() + 1) * sizeof(void *)))
        return 1;
    std::vector<const fe *> hp(polys.size() + 1, nullptr);
    for (size_t i = 0; i < polys.size(); i++) hp[i] = polys[i]->buf.as<fe>();
    CU(cudaMemcpyAsync(X.po_ptrs.p, hp.data(), hp.size() * sizeof(void *), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(X.ast_code.p, code, n_code * sizeof(AstInstr), cudaMemcpyHostToDevice, s));
    if (n_consts) {
        CU(cudaMemcpyAsync(X.ast_consts.p, consts, n_consts * sizeof(fe), cudaMemcpyHostToDevice, s));
        if (repr == H2_REPR_CANONICAL) LAUNCH(convert_kernel<P>, blocks_for(n_consts, 64), 64, 0, s, X.ast_consts.as<fe>(), (uint64_t)n_consts, 1);
    }
    AstArgs A;
    A.polys = X.po_ptrs.as<const fe *>(); A.code = X.ast_code.as<AstInstr>(); A.n_code = (uint32_t)n_code; A.consts = X.ast_consts.as<fe>();
    A.tw = nullptr; A.lin_base = fe_one<P>(); A.log_n = log_n; A.out = out->buf.as<fe>();
    if (has_linear) {
        if (get_twiddles<P>(out->field, host_to_mont<P>(omega, repr), log_n, s, &A.tw)) return 1;
        A.lin_base = host_to_mont<P>(lin_base, repr);
    }
    LAUNCH(ast_eval_kernel<P>, blocks_for(n, 128), 128, 0, s, A);
    return scratch_release(s);
}
extern "C" int h2_poly_eval_ast(uint64_t out, const uint64_t *polys, size_t n_polys, uint32_t log_n, const uint32_t *code, size_t n_code,
                                const void *consts, size_t n_consts, const void *omega, const void *lin_base, int repr) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    PolyBuf *o = find_poly(out);
    if (!o) return fail("h2_poly_eval_ast: unknown output handle");
    if (log_n > 30 || o->len < ((size_t)1 << log_n)) return fail("h2_poly_eval_ast: the output holds fewer than 2^log_n elements");
    if (n_code == 0 || n_code > (1u << 20)) return fail("h2_poly_eval_ast: empty or oversized program");
    std::vector<PolyBuf *> ps(n_polys);
    for (size_t i = 0; i < n_polys; i++) {
        ps[i] = find_poly(polys[i]);
        if (!ps[i]) return fail("h2_poly_eval_ast: unknown polynomial handle");
        if (ps[i]->field != o->field) return fail("h2_poly_eval_ast: the polynomials live in different fields");
        if (ps[i]->len < ((size_t)1 << log_n)) return fail("h2_poly_eval_ast: a polynomial holds fewer than 2^log_n elements");
        if (ps[i] == o) return fail("h2_poly_eval_ast: the output cannot be one of the operands (rotated reads)");
    }
    const AstInstr *prog = reinterpret_cast<const AstInstr *>(code);
    int depth = 0;
    bool has_linear = false;
    for (size_t pc = 0; pc < n_code; pc++) {
        const AstInstr &in = prog[pc];
        switch (in.op) {
        case AST_POLY: if (in.arg >= n_polys) return fail("h2_poly_eval_ast: polynomial index out of range"); depth++; break;
        case AST_LINEAR: has_linear = true;   /* fall through */
        case AST_CONST: if (in.arg >= n_consts) return fail("h2_poly_eval_ast: constant index out of range"); depth++; break;
        case AST_ADD: case AST_MUL: if (depth < 2) return fail("h2_poly_eval_ast: operand stack underflow"); depth--; break;
        case AST_SCALE: if (in.arg >= n_consts) return fail("h2_poly_eval_ast: constant index out of range");   /* fall through */
        case AST_NEG: if (depth < 1) return fail("h2_poly_eval_ast: operand stack underflow"); break;
        default: return fail("h2_poly_eval_ast: unknown opcode");
        }
        if (depth > H2_AST_STACK) return fail("h2_poly_eval_ast: expression deeper than the operand stack (24)");
    }
    if (depth != 1) return fail("h2_poly_eval_ast: the program must leave exactly one value");
    if (has_linear && (!omega || !lin_base)) return fail("h2_poly_eval_ast: a LinearTerm needs omega and the coset generator");
    if (o->field == H2_FIELD_FP) return ast_run<FpParams>(o, ps, log_n, prog, n_code, consts, n_consts, omega, lin_base, repr, has_linear);
    return ast_run<FqParams>(o, ps, log_n, prog, n_code, consts, n_consts, omega, lin_base, repr, has_linear);
}
// ff::BatchInvert on the first n elements of a resident polynomial, in place (zeros stay zero)
extern "C" int h2_poly_batch_invert(uint64_t poly, size_t n) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    PolyBuf *a = find_poly(poly);
    if (!a) return fail("h2_poly_batch_invert: unknown polynomial handle");
    if (a->len < n) return fail("h2_poly_batch_invert: the polynomial holds fewer than n elements");
    if (n == 0) return 0;
    cudaStream_t s = g_ctx.stream;
    if (scratch_acquire(s)) return 1;
    const uint32_t nb = blocks_for((n + 15) / 16, 64);
    if (a->field == H2_FIELD_FP) LAUNCH(poly_batch_invert_kernel<FpParams>, nb, 64, 0, s, a->buf.as<fe>(), (uint64_t)n);
    else LAUNCH(poly_batch_invert_kernel<FqParams>, nb, 64, 0, s, a->buf.as<fe>(), (uint64_t)n);
    return scratch_release(s);
}
// dst[0] = init, dst[i] = dst[i - 1] * src[i - 1] for i < n: the running product of plonk/permutation/prover.rs:150-156
template <class P> static int grand_product_run(PolyBuf *d, PolyBuf *a, size_t n, const void *init, int repr) {
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    std::vector<uint64_t> m{(uint64_t)n}, off{0};
    while (m.back() > H2_POLY_CHUNK) { off.push_back(off.back() + (m.size() > 1 ? m.back() : 0)); m.push_back((m.back() + H2_POLY_CHUNK - 1) / H2_POLY_CHUNK); }
    const size_t L = m.size() - 1;
    uint64_t total = 1;
    for (size_t l = 1; l <= L; l++) total += m[l];
    if (scratch_acquire(s)) return 1;
    if (X.po_lvl.ensure(total * sizeof(fe)) || X.po_q.ensure(total * sizeof(fe))) return 1;
    fe *lvl = X.po_lvl.as<fe>(), *ex = X.po_q.as<fe>();
    const fe *src = a->buf.as<fe>();
    const fe in0 = host_to_mont<P>(init, repr);
    for (size_t l = 0; l < L; l++)
        LAUNCH(poly_product_up_kernel<P>, blocks_for(m[l + 1], 128), 128, 0, s, l == 0 ? src : (const fe *)(lvl + off[l]), m[l], lvl + off[l + 1], m[l + 1]);
    for (size_t l = L + 1; l-- > 0;) {
        const uint64_t chunks = (m[l] + H2_POLY_CHUNK - 1) / H2_POLY_CHUNK;
        LAUNCH(poly_product_down_kernel<P>, blocks_for(chunks, 128), 128, 0, s, l == 0 ? src : (const fe *)(lvl + off[l]), m[l],
               l == L ? (const fe *)nullptr : (const fe *)(ex + off[l + 1]), in0, l == 0 ? d->buf.as<fe>() : ex + off[l], chunks);
    }
    return scratch_release(s);
}
extern "C" int h2_poly_running_product(uint64_t dst, uint64_t src, size_t n, const void *init, int repr) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    PolyBuf *d = find_poly(dst), *a = find_poly(src);
    if (!d || !a) return fail("h2_poly_running_product: unknown polynomial handle");
    if (d == a) return fail("h2_poly_running_product: the product cannot overwrite its factors");
    if (d->field != a->field) return fail("h2_poly_running_product: the polynomials live in different fields");
    if (a->len < n || d->len < n) return fail("h2_poly_running_product: a polynomial holds fewer than n elements");
    if (n == 0) return 0;
    if (a->field == H2_FIELD_FP) return grand_product_run<FpParams>(d, a, n, init, repr);
    return grand_product_run<FqParams>(d, a, n, init, repr);
}
// divide_by_vanishing_poly on a resident extended-domain polynomial; t_evals: t_len = 2^(ext_k - k) host elements
extern "C" int h2_poly_divide_by_vanishing(uint64_t poly, uint32_t ext_k, const void *t_evals, uint32_t t_len, int repr) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    PolyBuf *a = find_poly(poly);
    if (!a) return fail("h2_poly_divide_by_vanishing: unknown polynomial handle");
    if (ext_k > 30 || a->len < ((size_t)1 << ext_k)) return fail("h2_poly_divide_by_vanishing: the polynomial holds fewer than 2^ext_k elements");
    if (t_len == 0 || (t_len & (t_len - 1)) || t_len > (1u << ext_k)) return fail("h2_poly_divide_by_vanishing: t_len must be a power of two <= 2^ext_k");
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    if (scratch_acquire(s)) return 1;
    if (X.po_pts.ensure((size_t)t_len * sizeof(fe))) return 1;
    CU(cudaMemcpyAsync(X.po_pts.p, t_evals, (size_t)t_len * sizeof(fe), cudaMemcpyHostToDevice, s));
    if (repr == H2_REPR_CANONICAL && convert_field(a->field, X.po_pts.as<fe>(), t_len, 1, s)) return 1;
    const uint64_t n = 1ull << ext_k;
    if (a->field == H2_FIELD_FP) LAUNCH(poly_vanish_div_kernel<FpParams>, blocks_for(n, 256), 256, 0, s, a->buf.as<fe>(), n, (const fe *)X.po_pts.as<fe>(), t_len - 1);
    else LAUNCH(poly_vanish_div_kernel<FqParams>, blocks_for(n, 256), 256, 0, s, a->buf.as<fe>(), n, (const fe *)X.po_pts.as<fe>(), t_len - 1);
    return scratch_release(s);
}
extern "C" int h2_poly_eval(const uint64_t *polys, size_t batch, size_t n, const void *points, int repr, void *out) {
    if (n == 0) { memset(out, 0, batch * 32); return 0; }            // the empty sum (fold over nothing, arithmetic.rs:300-302)
    return polyops_dispatch(0, polys, nullptr, batch, n, points, repr, out, "h2_poly_eval");
}
extern "C" int h2_poly_inner_product(const uint64_t *a, const uint64_t *b, size_t batch, size_t n, int repr, void *out) {
    if (n == 0) { memset(out, 0, batch * 32); return 0; }
    return polyops_dispatch(1, a, b, batch, n, nullptr, repr, out, "h2_poly_inner_product");
}
extern "C" int h2_poly_kate_division(const uint64_t *dst, const uint64_t *src, size_t batch, size_t n, const void *points, int repr) {
    if (n == 0) return fail("h2_poly_kate_division: empty polynomial (the reference underflows a.len() - 1, arithmetic.rs:329)");
    if (n == 1) return 0;                                             // quotient of a constant: no coefficients
    return polyops_dispatch(2, src, dst, batch, n, points, repr, nullptr, "h2_poly_kate_division");
}

// ------------------------------------------------------------------------------------------------
static void ipa_free(IpaSession *q) {   // back to the pool (the caller has synchronised the stream)
    if (g_ctx.ipa_pool.size() < 2) { g_ctx.ipa_pool.push_back(q); return; }
    q->p.release(); q->b.release(); q->s.release(); q->scal.release(); q->out.release(); delete q;
}
static IpaState ipa_state(IpaSession *q) {
    IpaState S;
    S.p = q->p.as<fe>(); S.b = q->b.as<fe>(); S.s = q->s.as<fe>(); S.scal = q->scal.as<fe>(); S.n = 1ull << q->k;
    return S;
}
template <class PS> static int ipa_begin_impl(IpaSession *q, const void *p_prime, const void *x3, int repr, cudaStream_t s) {
    Context &X = g_ctx;
    const uint64_t n = 1ull << q->k;
    if (q->p.ensure(n * sizeof(fe)) || q->b.ensure(n * sizeof(fe)) || q->s.ensure(n * sizeof(fe)) || q->scal.ensure(2 * (n + 2) * sizeof(fe)) ||
        q->out.ensure(2 * sizeof(jacobian)) || X.pow2.ensure(64 * sizeof(fe)))
        return 1;
    CU(cudaMemcpyAsync(q->p.p, p_prime, n * sizeof(fe), cudaMemcpyHostToDevice, s));
    IpaState S = ipa_state(q);
    LAUNCH(ipa_init_kernel<PS>, blocks_for(n, 256), 256, 0, s, S, repr == H2_REPR_MONTGOMERY);
    // b_t = x3^t (prover.rs:86-93) with the NTT twiddle generator
    fe x = host_to_mont<PS>(x3, repr);
    LAUNCH(twiddle_pow2_kernel<PS>, 1, 32, 0, s, X.pow2.as<fe>(), x, q->k + 1);
    LAUNCH(twiddle_fill_kernel<PS>, blocks_for((n + 31) / 32, 128), 128, 0, s, S.b, X.pow2.as<fe>(), n);
    return 0;
}
extern "C" int h2_ipa_begin(uint64_t bases_handle, uint32_t k, const void *p_prime, const void *x3, int repr, uint64_t *session) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    auto it = g_ctx.bases.find(bases_handle);
    if (it == g_ctx.bases.end()) return fail("h2_ipa_begin: unknown bases handle");
    BaseSet *b = it->second;
    if (k == 0 || k > 28) return fail("h2_ipa_begin: k out of range");
    if (b->n != (1ull << k) + 2) return fail("h2_ipa_begin: the base set must hold g[0..2^k) || w || u");
    if (!b->table.p) return fail("h2_ipa_begin: the base set has no window table (register with H2_BASES_PRECOMPUTE)");
    IpaSession *q;
    if (!g_ctx.ipa_pool.empty()) { q = g_ctx.ipa_pool.back(); g_ctx.ipa_pool.pop_back(); }
    else q = new IpaSession();
    q->bases = bases_handle; q->k = k; q->round = 0; q->folded = 1;
    cudaStream_t s = g_ctx.stream;
    if (scratch_acquire(s)) { ipa_free(q); return 1; }   // pow2 is shared scratch
    int rc = b->curve == H2_CURVE_PALLAS ? ipa_begin_impl<FqParams>(q, p_prime, x3, repr, s) : ipa_begin_impl<FpParams>(q, p_prime, x3, repr, s);
    if (rc) { ipa_free(q); return 1; }
    if (scratch_release(s)) { ipa_free(q); return 1; }
    cudaError_t e = cudaStreamSynchronize(s);   // p_prime may be pageable host memory
    if (e != cudaSuccess) { ipa_free(q); return fail(std::string("h2_ipa_begin: ") + cudaGetErrorString(e)); }
    uint64_t h = g_ctx.next_handle++;
    g_ctx.ipa[h] = q;
    *session = h;
    return 0;
}
template <class PS> static int ipa_round_impl(IpaSession *q, BaseSet *b, const void *z, const void *l_rand, const void *r_rand, int repr, cudaStream_t s) {
    const uint64_t n = 1ull << q->k;
    const uint32_t bit = q->k - 1 - q->round;
    IpaState S = ipa_state(q);
    LAUNCH(ipa_prep_kernel<PS>, blocks_for(n, 256), 256, 0, s, S, bit);
    LAUNCH(ipa_inner_kernel<PS>, 1, 512, 0, s, S, bit, host_to_mont<PS>(z, repr), host_to_mont<PS>(l_rand, repr), host_to_mont<PS>(r_rand, repr));
    uint32_t tc, tmode;
    const affine *tbl = fixed_table(b, &tc, &tmode);
    return msm_dispatch(b->curve, S.scal, 1, tbl, n + 2, tc, q->out.as<jacobian>(), repr == H2_REPR_CANONICAL, s, tmode, b->n, nullptr, 2);
}
extern "C" int h2_ipa_round(uint64_t session, const void *z, const void *l_rand, const void *r_rand, int repr, void *out_lr_xyz) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    auto it = g_ctx.ipa.find(session);
    if (it == g_ctx.ipa.end()) return fail("h2_ipa_round: unknown session");
    IpaSession *q = it->second;
    auto ib = g_ctx.bases.find(q->bases);
    if (ib == g_ctx.bases.end()) return fail("h2_ipa_round: the session's base set was released");
    if (q->round >= q->k) return fail("h2_ipa_round: all k rounds are done");
    if (!q->folded) return fail("h2_ipa_round: h2_ipa_fold must follow each round");
    BaseSet *b = ib->second;
    cudaStream_t s = g_ctx.stream;
    if (scratch_acquire(s)) return 1;
    int rc = b->curve == H2_CURVE_PALLAS ? ipa_round_impl<FqParams>(q, b, z, l_rand, r_rand, repr, s) : ipa_round_impl<FpParams>(q, b, z, l_rand, r_rand, repr, s);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out_lr_xyz, q->out.p, 2 * sizeof(jacobian), cudaMemcpyDeviceToHost, s));
    if (scratch_release(s)) return 1;
    CU(cudaStreamSynchronize(s));
    q->folded = 0;
    return 0;
}
extern "C" int h2_ipa_fold(uint64_t session, const void *u, const void *u_inv, int repr) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    auto it = g_ctx.ipa.find(session);
    if (it == g_ctx.ipa.end()) return fail("h2_ipa_fold: unknown session");
    IpaSession *q = it->second;
    auto ib = g_ctx.bases.find(q->bases);
    if (ib == g_ctx.bases.end()) return fail("h2_ipa_fold: the session's base set was released");
    if (q->folded) return fail("h2_ipa_fold: no round to fold");
    const uint64_t n = 1ull << q->k;
    const uint32_t bit = q->k - 1 - q->round;
    cudaStream_t s = g_ctx.stream;
    IpaState S = ipa_state(q);
    if (ib->second->curve == H2_CURVE_PALLAS) LAUNCH(ipa_fold_kernel<FqParams>, blocks_for(n, 256), 256, 0, s, S, bit, host_to_mont<FqParams>(u, repr), host_to_mont<FqParams>(u_inv, repr));
    else LAUNCH(ipa_fold_kernel<FpParams>, blocks_for(n, 256), 256, 0, s, S, bit, host_to_mont<FpParams>(u, repr), host_to_mont<FpParams>(u_inv, repr));
    q->round++; q->folded = 1;   // asynchronous: the next round (or finish) is ordered behind it on the stream
    return 0;
}
extern "C" int h2_ipa_finish(uint64_t session, int repr, void *out_c_b) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    auto it = g_ctx.ipa.find(session);
    if (it == g_ctx.ipa.end()) return fail("h2_ipa_finish: unknown session");
    IpaSession *q = it->second;
    auto ib = g_ctx.bases.find(q->bases);
    int rc = 0;
    cudaStream_t s = g_ctx.stream;
    if (out_c_b) {
        if (ib == g_ctx.bases.end()) rc = fail("h2_ipa_finish: the session's base set was released");
        else if (q->round != q->k || !q->folded) rc = fail("h2_ipa_finish: the k rounds are not complete");
        else {
            IpaState S = ipa_state(q);
            fe *out = q->scal.as<fe>();
            if (ib->second->curve == H2_CURVE_PALLAS) ipa_result_kernel<FqParams><<<1, 32, 0, s>>>(S, repr == H2_REPR_CANONICAL, out);
            else ipa_result_kernel<FpParams><<<1, 32, 0, s>>>(S, repr == H2_REPR_CANONICAL, out);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            cudaError_t e = cudaMemcpyAsync(out_c_b, out, 2 * sizeof(fe), cudaMemcpyDeviceToHost, s);
            if (e != cudaSuccess) rc = fail(std::string("h2_ipa_finish: ") + cudaGetErrorString(e));
        }
    }
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess && !rc) rc = fail(std::string("h2_ipa_finish: ") + cudaGetErrorString(e));
    ipa_free(q);
    g_ctx.ipa.erase(it);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// Device-resident polynomials: the transforms and commits of the quotient
// pipeline without a PCIe round trip per call.  Data is kept in Montgomery form; every buffer has one spare
// slot so that a commit can append the blind.
// ------------------------------------------------------------------------------------------------
static PolyBuf *find_poly(uint64_t h) {
    auto it = g_ctx.polys.find(h);
    return it == g_ctx.polys.end() ? nullptr : it->second;
}
extern "C" int h2_poly_alloc(int field, size_t len, uint64_t *poly) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    if (field != H2_FIELD_FP && field != H2_FIELD_FQ) return fail("unknown field id");
    PolyBuf *b = new PolyBuf();
    b->field = field; b->len = len;
    if (b->buf.ensure((len + 1) * sizeof(fe))) { delete b; return 1; }
    uint64_t h = g_ctx.next_handle++;
    g_ctx.polys[h] = b;
    *poly = h;
    return 0;
}
extern "C" int h2_poly_free(uint64_t poly) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_ctx.polys.find(poly);
    if (it == g_ctx.polys.end()) return fail("h2_poly_free: unknown handle");
    cudaSetDevice(g_ctx.device);
    cudaStreamSynchronize(g_ctx.stream);
    it->second->buf.release();
    delete it->second;
    g_ctx.polys.erase(it);
    return 0;
}
static int convert_field(int field, fe *d, size_t n, int to_mont, cudaStream_t s) {
    if (n == 0) return 0;
    if (field == H2_FIELD_FP) LAUNCH(convert_kernel<FpParams>, blocks_for(n, 256), 256, 0, s, d, (uint64_t)n, to_mont);
    else LAUNCH(convert_kernel<FqParams>, blocks_for(n, 256), 256, 0, s, d, (uint64_t)n, to_mont);
    return 0;
}
extern "C" int h2_poly_upload(uint64_t poly, const void *src, size_t len, int repr) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    PolyBuf *b = find_poly(poly);
    if (!b) return fail("h2_poly_upload: unknown handle");
    if (len > b->len) return fail("h2_poly_upload: more elements than the polynomial holds");
    cudaStream_t s = g_ctx.stream;
    CU(cudaMemcpyAsync(b->buf.p, src, len * sizeof(fe), cudaMemcpyHostToDevice, s));
    if (repr == H2_REPR_CANONICAL && convert_field(b->field, b->buf.as<fe>(), len, 1, s)) return 1;
    CU(cudaStreamSynchronize(s));      // src may be pageable
    return 0;
}
extern "C" int h2_poly_download(uint64_t poly, void *dst, size_t len, int repr) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    PolyBuf *b = find_poly(poly);
    if (!b) return fail("h2_poly_download: unknown handle");
    if (len > b->len) return fail("h2_poly_download: more elements than the polynomial holds");
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    const fe *from = b->buf.as<fe>();
    if (repr == H2_REPR_CANONICAL) {   // convert a copy: the resident data stays in Montgomery form
        if (scratch_acquire(s)) return 1;
        if (X.ntt_out.ensure(len * sizeof(fe))) return 1;
        CU(cudaMemcpyAsync(X.ntt_out.p, from, len * sizeof(fe), cudaMemcpyDeviceToDevice, s));
        if (convert_field(b->field, X.ntt_out.as<fe>(), len, 0, s)) return 1;
        from = X.ntt_out.as<fe>();
    }
    CU(cudaMemcpyAsync(dst, from, len * sizeof(fe), cudaMemcpyDeviceToHost, s));
    if (repr == H2_REPR_CANONICAL && scratch_release(s)) return 1;
    CU(cudaStreamSynchronize(s));
    return 0;
}
// mode as in ntt_host: 1 = inverse transform with divisor, 2 = coeff_to_extended, 3 = extended_to_coeff
template <class P>
static int poly_transform(PolyBuf *dst, PolyBuf *src, int mode, uint32_t in_log_n, uint32_t log_n, const void *omega, const void *zeta,
                          const void *divisor, size_t out_len, int repr) {
    cudaStream_t s = g_ctx.stream;
    if (scratch_acquire(s)) return 1;
    fe w = host_to_mont<P>(omega, repr), z, d;
    if (zeta) z = host_to_mont<P>(zeta, repr);
    if (divisor) d = host_to_mont<P>(divisor, repr);
    NttScales sc = make_scales<P>(H2_REPR_MONTGOMERY, mode == 2 ? &z : nullptr, (mode == 1 || mode == 3) ? &d : nullptr, mode == 3 ? &z : nullptr);
    if (ntt_run<P>(src->field, src->buf.as<fe>(), in_log_n, dst->buf.as<fe>(), log_n, w, sc, out_len, s)) return 1;
    return scratch_release(s);       // asynchronous: later calls are ordered behind it on the stream
}
static int poly_transform_dispatch(uint64_t dst, uint64_t src, int mode, uint32_t in_log_n, uint32_t log_n, const void *omega, const void *zeta,
                                   const void *divisor, size_t out_len, int repr) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    PolyBuf *d = find_poly(dst), *a = find_poly(src);
    if (!d || !a) return fail("resident transform: unknown polynomial handle");
    if (d->field != a->field) return fail("resident transform: the polynomials live in different fields");
    if (log_n > 30 || in_log_n > log_n) return fail("ntt: bad sizes");
    if (a->len < ((size_t)1 << in_log_n)) return fail("resident transform: the source holds fewer than 2^k elements");
    if (out_len > ((size_t)1 << log_n)) out_len = (size_t)1 << log_n;
    if (d->len < out_len) return fail("resident transform: the destination is too short");
    if (d == a && out_len != ((size_t)1 << log_n)) return fail("resident transform: in place needs out_len == 2^log_n");
    if (d == a && in_log_n != log_n) return fail("resident transform: in place needs equal input and output sizes");
    if (a->field == H2_FIELD_FP) return poly_transform<FpParams>(d, a, mode, in_log_n, log_n, omega, zeta, divisor, out_len, repr);
    return poly_transform<FqParams>(d, a, mode, in_log_n, log_n, omega, zeta, divisor, out_len, repr);
}
extern "C" int h2_poly_lagrange_to_coeff(uint64_t dst, uint64_t src, uint32_t k, const void *omega_inv, const void *divisor, int repr) {
    return poly_transform_dispatch(dst, src, 1, k, k, omega_inv, nullptr, divisor, (size_t)1 << k, repr);
}
extern "C" int h2_poly_coeff_to_extended(uint64_t dst, uint64_t src, uint32_t k, uint32_t ext_k, const void *zeta, const void *ext_omega, int repr) {
    return poly_transform_dispatch(dst, src, 2, k, ext_k, ext_omega, zeta, nullptr, (size_t)1 << ext_k, repr);
}
extern "C" int h2_poly_extended_to_coeff(uint64_t dst, uint64_t src, uint32_t ext_k, const void *ext_omega_inv, const void *ext_divisor,
                                         const void *zeta, size_t out_len, int repr) {
    return poly_transform_dispatch(dst, src, 3, ext_k, ext_k, ext_omega_inv, zeta, ext_divisor, out_len, repr);
}
// commit(poly, blind) = <poly[0..n), bases[0..n)> + blind * bases[n] for `batch` resident polynomials in one pass
extern "C" int h2_msm_registered_polys(uint64_t bases_handle, const uint64_t *polys, size_t batch, size_t n, const void *extra_scalars, int repr,
                                       void *out_xyz) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    auto it = g_ctx.bases.find(bases_handle);
    if (it == g_ctx.bases.end()) return fail("h2_msm_registered_polys: unknown bases handle");
    BaseSet *b = it->second;
    if (batch == 0) return 0;
    if (batch > 64) return fail("h2_msm_registered_polys: batch > 64");
    if (batch > 1 && !b->table.p) return fail("h2_msm_registered_polys: a batch needs a base set with a window table (H2_BASES_PRECOMPUTE)");
    const size_t total = n + (extra_scalars ? 1 : 0);
    if (total > b->n) return fail("h2_msm_registered_polys: more scalars than registered bases");
    const int scalar_field = b->curve == H2_CURVE_PALLAS ? H2_FIELD_FQ : H2_FIELD_FP;
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    if (scratch_acquire(s)) return 1;
    if (X.scal_in.ensure(batch * total * sizeof(fe)) || X.result.ensure(batch * sizeof(jacobian)) || X.misc.ensure(batch * sizeof(fe) + 64)) return 1;
    fe *d = X.scal_in.as<fe>();
    if (extra_scalars) {   // the blinds: Montgomery form like the resident data
        CU(cudaMemcpyAsync(X.misc.p, extra_scalars, batch * sizeof(fe), cudaMemcpyHostToDevice, s));
        if (repr == H2_REPR_CANONICAL && convert_field(scalar_field, X.misc.as<fe>(), batch, 1, s)) return 1;
    }
    for (size_t j = 0; j < batch; j++) {
        PolyBuf *q = find_poly(polys[j]);
        if (!q) return fail("h2_msm_registered_polys: unknown polynomial handle");
        if (q->field != scalar_field) return fail("h2_msm_registered_polys: the polynomial is not over the curve's scalar field");
        if (q->len < n) return fail("h2_msm_registered_polys: the polynomial holds fewer than n elements");
        CU(cudaMemcpyAsync(d + j * total, q->buf.p, n * sizeof(fe), cudaMemcpyDeviceToDevice, s));
        if (extra_scalars) CU(cudaMemcpyAsync(d + j * total + n, X.misc.as<fe>() + j, sizeof(fe), cudaMemcpyDeviceToDevice, s));
    }
    int rc;
    uint32_t tc = 0, tmode = 0;
    const affine *tbl = b->table.p ? fixed_table(b, &tc, &tmode) : nullptr;
    if (tbl) rc = msm_dispatch(b->curve, d, 1, tbl, total, tc, X.result.as<jacobian>(), repr == H2_REPR_CANONICAL, s, tmode, b->n,
                               nullptr, (uint32_t)batch);
    else rc = msm_dispatch(b->curve, d, 1, b->buf.as<affine>(), total, 0, X.result.as<jacobian>(), repr == H2_REPR_CANONICAL, s);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out_xyz, X.result.p, batch * sizeof(jacobian), cudaMemcpyDeviceToHost, s));
    if (scratch_release(s)) return 1;
    CU(cudaStreamSynchronize(s));
    return 0;
}

extern "C" int h2_dev_gen_points(int curve, uint64_t seed, uint64_t first, size_t n, void *d_out, void *stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return 0;
    if (curve == H2_CURVE_PALLAS) LAUNCH(gen_points_kernel<FpParams>, blocks_for(n, 128), 128, 0, s, (affine *)d_out, seed, first, (uint64_t)n);
    else if (curve == H2_CURVE_VESTA) LAUNCH(gen_points_kernel<FqParams>, blocks_for(n, 128), 128, 0, s, (affine *)d_out, seed, first, (uint64_t)n);
    else return fail("unknown curve id");
    return 0;
}
extern "C" int h2_dev_convert(int field, void *d_a, size_t n, int to_montgomery, void *stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return 0;
    if (field == H2_FIELD_FP) LAUNCH(convert_kernel<FpParams>, blocks_for(n, 256), 256, 0, s, (fe *)d_a, (uint64_t)n, to_montgomery);
    else if (field == H2_FIELD_FQ) LAUNCH(convert_kernel<FqParams>, blocks_for(n, 256), 256, 0, s, (fe *)d_a, (uint64_t)n, to_montgomery);
    else return fail("unknown field id");
    return 0;
}
extern "C" int h2_test_field_op(int field, int op, const void *a, const void *b, size_t n, void *out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    if (scratch_acquire(s)) return 1;
    if (X.misc.ensure(3 * n * sizeof(fe) + 64)) return 1;
    fe *da = X.misc.as<fe>(), *db = da + n, *dout = db + n;
    CU(cudaMemcpyAsync(da, a, n * sizeof(fe), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(db, b, n * sizeof(fe), cudaMemcpyHostToDevice, s));
    if (field == H2_FIELD_FP) LAUNCH(test_field_kernel<FpParams>, blocks_for(n, 128), 128, 0, s, da, db, dout, (uint64_t)n, op);
    else LAUNCH(test_field_kernel<FqParams>, blocks_for(n, 128), 128, 0, s, da, db, dout, (uint64_t)n, op);
    CU(cudaMemcpyAsync(out, dout, n * sizeof(fe), cudaMemcpyDeviceToHost, s));
    if (scratch_release(s)) return 1;
    CU(cudaStreamSynchronize(s));
    return 0;
}
extern "C" int h2_test_curve_op(int curve, int op, const void *a_xy, const void *b_xy, size_t n, void *out_xy) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    if (scratch_acquire(s)) return 1;
    if (X.misc.ensure(3 * n * sizeof(affine) + 64)) return 1;
    affine *da = X.misc.as<affine>(), *db = da + n, *dout = db + n;
    CU(cudaMemcpyAsync(da, a_xy, n * sizeof(affine), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(db, b_xy, n * sizeof(affine), cudaMemcpyHostToDevice, s));
    if (curve == H2_CURVE_PALLAS) LAUNCH(test_curve_kernel<FpParams>, blocks_for(n, 64), 64, 0, s, da, db, dout, (uint64_t)n, op);
    else LAUNCH(test_curve_kernel<FqParams>, blocks_for(n, 64), 64, 0, s, da, db, dout, (uint64_t)n, op);
    CU(cudaMemcpyAsync(out_xy, dout, n * sizeof(affine), cudaMemcpyDeviceToHost, s));
    if (scratch_release(s)) return 1;
    CU(cudaStreamSynchronize(s));
    return 0;
}
extern "C" int h2_bench_field_mul(int field, uint32_t threads_per_block, uint32_t blocks, uint32_t iters, float *ms) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    if (scratch_acquire(s)) return 1;
    size_t threads = (size_t)threads_per_block * blocks;
    if (X.misc.ensure(threads * 4 * sizeof(fe))) return 1;
    CU(cudaMemsetAsync(X.misc.p, 0x11, threads * 4 * sizeof(fe), s));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; rep++) {   // first repetition warms up
        CU(cudaEventRecord(e0, s));
        // field | 0x100: the same loop with fe_sqr
        if (field == H2_FIELD_FP) LAUNCH((bench_mul_kernel<FpParams, false>), blocks, threads_per_block, 0, s, X.misc.as<fe>(), iters);
        else if (field == H2_FIELD_FQ) LAUNCH((bench_mul_kernel<FqParams, false>), blocks, threads_per_block, 0, s, X.misc.as<fe>(), iters);
        else if (field == (H2_FIELD_FP | 0x100)) LAUNCH((bench_mul_kernel<FpParams, true>), blocks, threads_per_block, 0, s, X.misc.as<fe>(), iters);
        else LAUNCH((bench_mul_kernel<FqParams, true>), blocks, threads_per_block, 0, s, X.misc.as<fe>(), iters);
        CU(cudaEventRecord(e1, s));
        CU(cudaStreamSynchronize(s));
    }
    CU(cudaEventElapsedTime(ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return scratch_release(s);
}

// mode: 0 dependent mul chain, 1 two chains, 2 four chains, 3 xyzz_double, 4 xyzz_add, 5 xyzz_add_mixed;
// one warp, `iters` iterations; *ms = elapsed.
extern "C" int h2_bench_latency(int mode, uint32_t iters, float *ms) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    if (scratch_acquire(s)) return 1;
    if (X.misc.ensure(32 * 4 * sizeof(fe))) return 1;
    CU(cudaMemsetAsync(X.misc.p, 0x11, 32 * 4 * sizeof(fe), s));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; rep++) {
        CU(cudaEventRecord(e0, s));
        LAUNCH(bench_latency_kernel<FpParams>, 1, 32, 0, s, X.misc.as<fe>(), iters, mode);
        CU(cudaEventRecord(e1, s));
        CU(cudaStreamSynchronize(s));
    }
    CU(cudaEventElapsedTime(ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return scratch_release(s);
}

// ------------------------------------------------------------------------------------------------
// per-kernel timing for the roofline leg of bench.py
// ------------------------------------------------------------------------------------------------
extern "C" int h2_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    cudaDeviceSynchronize();
    for (auto &sp : g_prof) { cudaEventDestroy(sp.e0); cudaEventDestroy(sp.e1); }
    g_prof.clear();
    g_prof_on = on != 0;
    return 0;
}
// kind 0 = msm_accum0_kernel, 1 = ntt_pass_kernel.  Returns summed device time and launch count.
extern "C" int h2_profile_read(int kind, float *total_ms, uint32_t *launches) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    CU(cudaDeviceSynchronize());
    float tot = 0; uint32_t cnt = 0;
    for (auto &sp : g_prof) {
        if (sp.kind != kind) continue;
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, sp.e0, sp.e1));
        tot += ms; cnt++;
    }
    *total_ms = tot; *launches = cnt;
    return 0;
}
