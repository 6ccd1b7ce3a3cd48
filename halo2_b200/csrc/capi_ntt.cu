// C ABI of the engine, part 3 of 5: the NTT pipeline (ntt.cuh), the domain transforms, device-resident polynomials.
#include "util_kernels.cuh"
#include "ntt.cuh"

// ------------------------------------------------------------------------------------------------
// NTT pipeline
// ------------------------------------------------------------------------------------------------
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
        uint32_t smem = ntt_smem_bytes(sp[i], logc[i]) + ntt_twc_bytes(sp[i], logc[i], i == passes - 1);
        static bool smem_optin = false;      // per instantiation <P>: a single-CTA transform of 2^10 elements wants 64 KiB
        if (!smem_optin) {
            CU(cudaFuncSetAttribute(ntt_pass_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            CU(cudaFuncSetAttribute(ntt_pass_tma_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            smem_optin = true;
        }
        prof_begin(PROF_NTT_PASS, s);
        if (X.ntt_tma && NttDense<P>::supported(A)) {
            // persistent CTAs (4 per SM by registers), double-buffered tiles on the bulk-copy engine
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, X.device);
            // every CTA walks the same number of tiles (+-1): grid = tiles / ceil(tiles / resident CTAs)
            const uint32_t slots = (uint32_t)sms * 4u, per = (tiles + slots - 1) / slots;
            const uint32_t grid = (tiles + per - 1) / per;
            LAUNCH(ntt_pass_tma_kernel<P>, grid, 128, ntt_tma_smem_bytes(sp[i], logc[i], i == passes - 1), s, A, tiles);
        } else
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
    if (upload_async(X.ntt_io.p, a_in, n_in * sizeof(fe), s)) return 1;
    if (ntt_run<P>(field, X.ntt_io.as<fe>(), in_log_n, X.ntt_out.as<fe>(), log_n, w, sc, out_len, s)) return 1;
    if (download_sync(out, X.ntt_out.p, out_len * sizeof(fe), s)) return 1;
    return scratch_release(s);
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
// test / bench hook: 1 = the bulk-copy (TMA) persistent pass kernel where it applies, 0 = the classic kernel (default: measured faster)
extern "C" int h2_test_set_ntt_tma(int on) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_ctx.ntt_tma = on ? 1u : 0u;
    return 0;
}
extern "C" int h2_ntt_clear_cache(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    cudaDeviceSynchronize();
    for (auto *t : g_ctx.twiddles) { t->buf.release(); delete t; }
    g_ctx.twiddles.clear();
    return 0;
}

int get_twiddles_any(int field, const fe &omega_mont, uint32_t log_n, cudaStream_t s, const fe **out) {
    if (field == H2_FIELD_FP) return get_twiddles<FpParams>(field, omega_mont, log_n, s, out);
    if (field == H2_FIELD_FQ) return get_twiddles<FqParams>(field, omega_mont, log_n, s, out);
    return fail("unknown field id");
}

// ------------------------------------------------------------------------------------------------
// Device-resident polynomials: the transforms and commits of the quotient
// pipeline without a PCIe round trip per call.  Data is kept in Montgomery form; every buffer has one spare
// slot so that a commit can append the blind.
// ------------------------------------------------------------------------------------------------
extern "C" int h2_poly_alloc(int field, size_t len, uint64_t *poly) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    if (field != H2_FIELD_FP && field != H2_FIELD_FQ) return fail("unknown field id");
    // A prover allocates and frees the same few sizes proof after proof: freed polynomials keep their device buffer in a small
    // pool, so that this is a memset on the stream instead of a cudaMalloc (and h2_poly_free no cudaFree + device sync).
    Context &X = g_ctx;
    PolyBuf *b = nullptr;
    for (size_t i = 0; i < X.poly_pool.size(); i++) {
        PolyBuf *c = X.poly_pool[i];
        if (c->buf.cap >= (len + 1) * sizeof(fe) && c->buf.cap <= (len + 1) * sizeof(fe) * 9 / 8 + 512) {
            b = c;
            X.poly_pool_bytes -= c->buf.cap;
            X.poly_pool.erase(X.poly_pool.begin() + i);
            break;
        }
    }
    if (!b) b = new PolyBuf();
    b->field = field; b->len = len;
    if (b->buf.ensure((len + 1) * sizeof(fe))) { delete b; return 1; }
    // zero-filled: a commit after a partial upload, or of a quotient shorter than the buffer, must not read stale memory
    if (cudaMemsetAsync(b->buf.p, 0, (len + 1) * sizeof(fe), X.stream) != cudaSuccess) { b->buf.release(); delete b; return fail("h2_poly_alloc: memset failed"); }
    uint64_t h = X.next_handle++;
    X.polys[h] = b;
    *poly = h;
    return 0;
}
extern "C" int h2_poly_free(uint64_t poly) {
    std::lock_guard<std::mutex> lk(g_mu);
    Context &X = g_ctx;
    auto it = X.polys.find(poly);
    if (it == X.polys.end()) return fail("h2_poly_free: unknown handle");
    PolyBuf *b = it->second;
    X.polys.erase(it);
    // every use of a resident polynomial is ordered on the context's stream, and so is its next owner's first write.
    // The pool is first-in first-out: when it is full the OLDEST buffers go (sizes an earlier workload left behind must not
    // pin the pool and push every later free onto the cudaFree + device-sync path: bench.py's replay ran 1.8 ms slower than
    // the same replay in a fresh process for exactly that reason)
    if (b->buf.cap > ((size_t)4 << 30)) {
        cudaSetDevice(X.device);
        cudaStreamSynchronize(X.stream);
        b->buf.release();
        delete b;
        return 0;
    }
    bool synced = false;
    while (!X.poly_pool.empty() && (X.poly_pool.size() >= 192 || X.poly_pool_bytes + b->buf.cap > ((size_t)4 << 30))) {
        PolyBuf *old = X.poly_pool.front();
        X.poly_pool.erase(X.poly_pool.begin());
        X.poly_pool_bytes -= old->buf.cap;
        if (!synced) { cudaSetDevice(X.device); cudaStreamSynchronize(X.stream); synced = true; }
        old->buf.release();
        delete old;
    }
    X.poly_pool.push_back(b);
    X.poly_pool_bytes += b->buf.cap;
    return 0;
}
int convert_field(int field, fe *d, size_t n, int to_mont, cudaStream_t s) {
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
    if (upload_async(b->buf.p, src, len * sizeof(fe), s)) return 1;
    if (repr == H2_REPR_CANONICAL && convert_field(b->field, b->buf.as<fe>(), len, 1, s)) return 1;
    CU(cudaStreamSynchronize(s));      // src may be pageable
    return 0;
}
// a[index] += delta: the one-coefficient corrections of the opening argument (poly/commitment/prover.rs:51 `s_poly[0] -= s_at_x3`,
// :78 `p_prime_poly[0] -= v`) on a resident polynomial
template <class P> __global__ void poly_add_at_kernel(fe *a, fe delta_mont) { fe_store(a, fe_add<P>(fe_load(a), delta_mont)); }
extern "C" int h2_poly_add_at(uint64_t poly, size_t index, const void *delta, int repr) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    PolyBuf *b = find_poly(poly);
    if (!b) return fail("h2_poly_add_at: unknown handle");
    if (index >= b->len) return fail("h2_poly_add_at: index out of range");
    cudaStream_t s = g_ctx.stream;
    if (b->field == H2_FIELD_FP) LAUNCH(poly_add_at_kernel<FpParams>, 1, 1, 0, s, b->buf.as<fe>() + index, host_to_mont<FpParams>(delta, repr));
    else LAUNCH(poly_add_at_kernel<FqParams>, 1, 1, 0, s, b->buf.as<fe>() + index, host_to_mont<FqParams>(delta, repr));
    return 0;
}
// dst[dst_off .. dst_off + len) = src[src_off .. src_off + len) on the device: the h(X) pieces (plonk/vanishing/prover.rs:95-100
// `h_poly.chunks_exact(n)`), or a copy of a column that an in-place step is about to overwrite
extern "C" int h2_poly_copy(uint64_t dst, size_t dst_off, uint64_t src, size_t src_off, size_t len) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    PolyBuf *d = find_poly(dst), *a = find_poly(src);
    if (!d || !a) return fail("h2_poly_copy: unknown handle");
    if (d->field != a->field) return fail("h2_poly_copy: the polynomials live in different fields");
    if (dst_off + len > d->len || src_off + len > a->len) return fail("h2_poly_copy: range out of bounds");
    if (d == a && !(dst_off + len <= src_off || src_off + len <= dst_off)) return fail("h2_poly_copy: overlapping ranges");
    if (len) CU(cudaMemcpyAsync(d->buf.as<fe>() + dst_off, a->buf.as<fe>() + src_off, len * sizeof(fe), cudaMemcpyDeviceToDevice, g_ctx.stream));
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
    if (download_sync(dst, from, len * sizeof(fe), s)) return 1;
    if (repr == H2_REPR_CANONICAL && scratch_release(s)) return 1;
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
