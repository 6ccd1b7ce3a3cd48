// C ABI of the engine (include/halo2_b200.h), part 1 of 5: context, device binding, settings, test hooks, utilities.
#include "util_kernels.cuh"

static thread_local std::string g_err;
int fail(const std::string &m) { g_err = m; return 1; }
uint64_t g_alloc_gen = 0;
Context g_ctxs[H2_MAX_DEVICES];
Context *g_cur = &g_ctxs[0];
std::mutex g_mu;
bool g_prof_on = false;
std::vector<ProfSpan> g_prof;
void prof_begin(int kind, cudaStream_t s) {
    if (!g_prof_on) return;
    ProfSpan sp; sp.kind = kind;
    cudaEventCreate(&sp.e0); cudaEventCreate(&sp.e1);
    cudaEventRecord(sp.e0, s);
    g_prof.push_back(sp);
}
void prof_end(cudaStream_t s) {
    if (!g_prof_on || g_prof.empty()) return;
    cudaEventRecord(g_prof.back().e1, s);
}
std::atomic<uint64_t> g_launches{0};

int require_ready() {
    if (!g_ctx.ready) return fail("h2_init has not been called (or failed): no CUDA device bound; there is no CPU fallback");
    CU(cudaSetDevice(g_ctx.device));
    return 0;
}
int scratch_acquire(cudaStream_t s) {
    if (g_ctx.have_last) CU(cudaStreamWaitEvent(s, g_ctx.last_use, 0));
    return 0;
}
int scratch_release(cudaStream_t s) {
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

