// C ABI of the engine, part 5 of 5: reductions and elementwise programs on resident polynomials (polyops.cuh, asteval.cuh).
#include "util_kernels.cuh"
#include "polyops.cuh"
#include "asteval.cuh"
#include "lookup.cuh"
#include "verifier.cuh"

// ------------------------------------------------------------------------------------------------
// eval_polynomial / compute_inner_product / kate_division on resident polynomials (polyops.cuh)
// ------------------------------------------------------------------------------------------------
// mode 0: eval (points: batch x 32 host), 1: inner product of a[i] and c[i], 2: kate division of a[i] by (X - point_i) into c[i]
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
    if (mode != 1 && n <= (1ull << 16) && n >= 2 && X.poly_cta) {
        // small polynomials: one CTA per polynomial does the whole reduction (polyops.cuh poly_eval_cta_kernel / poly_kate_cta_kernel)
        if (mode == 0) {
            fe *res = X.misc.as<fe>();
            LAUNCH(poly_eval_cta_kernel<P>, batch, H2_POLY_CTA, 0, s, d_a, (uint64_t)n, (const fe *)pts, res);
            if (repr == H2_REPR_CANONICAL) LAUNCH(convert_kernel<P>, blocks_for(batch, 64), 64, 0, s, res, (uint64_t)batch, 0);
            CU(cudaMemcpyAsync(out, res, batch * sizeof(fe), cudaMemcpyDeviceToHost, s));
            if (scratch_release(s)) return 1;
            CU(cudaStreamSynchronize(s));
            return 0;
        }
        LAUNCH(poly_kate_cta_kernel<P>, batch, H2_KATE_CTA, 0, s, d_a, (uint64_t)n, (const fe *)pts, (fe *const *)d_c);
        for (uint32_t b = 0; b < batch; b++) CU(cudaMemsetAsync(c[b]->buf.as<fe>() + (n - 1), 0, sizeof(fe), s));
        return scratch_release(s);
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
    // the quotient has n - 1 coefficients; slot n - 1 becomes the zero the reference pushes before committing n of them
    // (poly/multiopen/prover.rs: `kate_division(..); poly.push(ZERO)`)
    for (uint32_t b = 0; b < batch; b++) CU(cudaMemsetAsync(c[b]->buf.as<fe>() + (n - 1), 0, sizeof(fe), s));
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
        }
    }
    if (mode == 2)   // batch slices run concurrently: no quotient may be another slice's dividend, or be written twice
        for (size_t b = 0; b < batch; b++)
            for (size_t b2 = 0; b2 < batch; b2++) {
                if (c[b] == a[b2]) return fail(std::string(who) + ": the quotient cannot overwrite a dividend");
                if (b2 < b && c[b] == c[b2]) return fail(std::string(who) + ": a quotient handle appears twice");
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
        if (get_twiddles_any(out->field, host_to_mont<P>(omega, repr), log_n, s, &A.tw)) return 1;
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
// the lookup argument's permuted columns (lookup.cuh)
// ------------------------------------------------------------------------------------------------
static int lk_scan(uint32_t *d, uint64_t n, cudaStream_t s) {     // exclusive scan in place (kernels of msm.cuh)
    const uint64_t per_block = (uint64_t)H2_SCAN_BLOCK * H2_SCAN_ITEMS;
    const uint32_t nb = (uint32_t)((n + per_block - 1) / per_block);
    if (g_ctx.scan_blocks.ensure((size_t)nb * 4 + 16)) return 1;
    uint32_t *bs = g_ctx.scan_blocks.as<uint32_t>();
    LAUNCH(scan_block_sums_kernel, nb, H2_SCAN_BLOCK, 0, s, d, n, bs, (const uint32_t *)nullptr);
    LAUNCH(scan_single_block_kernel, 1, H2_SCAN_BLOCK, 0, s, bs, nb, (const uint32_t *)nullptr);
    LAUNCH(scan_apply_kernel, nb, H2_SCAN_BLOCK, 0, s, d, n, bs, (const uint32_t *)nullptr);
    return 0;
}
template <class P> static int lk_sort(fe *keys, uint64_t N, cudaStream_t s) {    // ascending bitonic sort of N = 2^m canonical keys
    const uint64_t BL = N < (1ull << H2_LK_BLOCK_LOG) ? N : (1ull << H2_LK_BLOCK_LOG);
    const uint32_t smem = (uint32_t)(BL * sizeof(fe)), thr = (uint32_t)(BL / 2 < 512 ? (BL / 2 ? BL / 2 : 1) : 512);
    LAUNCH(lk_bitonic_block_kernel, (uint32_t)(N / BL), thr, smem, s, keys, N, (uint64_t)2, 1u);
    for (uint64_t size = 2 * BL; size <= N; size <<= 1) {
        for (uint64_t stride = size / 2; stride >= BL; stride >>= 1)
            LAUNCH(lk_bitonic_global_kernel<P>, blocks_for(N / 2, 256), 256, 0, s, keys, N, size, stride);
        LAUNCH(lk_bitonic_block_kernel, (uint32_t)(N / BL), thr, smem, s, keys, N, size, 0u);
    }
    return 0;
}
template <class P> static int lookup_permute_run(PolyBuf *in, PolyBuf *tab, size_t u, PolyBuf *out_in, PolyBuf *out_tab) {
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    if (scratch_acquire(s)) return 1;
    uint64_t N = 2;
    while (N < u) N <<= 1;
    // u32 scratch: first flags | their scan (u + 1) | unconsumed flags | their scan (u + 1) | error word
    const size_t w = u + 1;
    if (X.lk_keys.ensure(2 * N * sizeof(fe)) || X.lk_left.ensure((u + 1) * sizeof(fe)) || X.lk_u32.ensure((4 * w + 4) * sizeof(uint32_t))) return 1;
    fe *ka = X.lk_keys.as<fe>(), *kt = ka + N, *left = X.lk_left.as<fe>();
    uint32_t *first = X.lk_u32.as<uint32_t>(), *first_scan = first + w, *unc = first_scan + w, *unc_scan = unc + w, *err = unc_scan + w;
    LAUNCH(lk_load_kernel<P>, blocks_for(N, 256), 256, 0, s, (const fe *)in->buf.as<fe>(), (uint64_t)u, ka, N);
    LAUNCH(lk_load_kernel<P>, blocks_for(N, 256), 256, 0, s, (const fe *)tab->buf.as<fe>(), (uint64_t)u, kt, N);
    if (lk_sort<P>(ka, N, s) || lk_sort<P>(kt, N, s)) return 1;
    CU(cudaMemsetAsync(first, 0, (4 * w + 4) * sizeof(uint32_t), s));
    LAUNCH(lk_fill_u32_kernel, blocks_for(u, 256), 256, 0, s, unc, (uint64_t)u, 1u);
    LAUNCH(lk_first_kernel<P>, blocks_for(u, 128), 128, 0, s, (const fe *)ka, (const fe *)kt, (uint64_t)u, first, unc, err, out_in->buf.as<fe>(),
           out_tab->buf.as<fe>());
    CU(cudaMemcpyAsync(first_scan, first, w * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(unc_scan, unc, w * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    if (lk_scan(first_scan, w, s) || lk_scan(unc_scan, w, s)) return 1;
    LAUNCH(lk_leftover_kernel<P>, blocks_for(u, 256), 256, 0, s, (const fe *)kt, (uint64_t)u, (const uint32_t *)unc, (const uint32_t *)unc_scan, left);
    LAUNCH(lk_fill_kernel<P>, blocks_for(u, 256), 256, 0, s, (uint64_t)u, (const uint32_t *)first, (const uint32_t *)first_scan, (const fe *)left,
           out_tab->buf.as<fe>());
    uint32_t h_err = 0;
    CU(cudaMemcpyAsync(&h_err, err, sizeof h_err, cudaMemcpyDeviceToHost, s));
    if (scratch_release(s)) return 1;
    CU(cudaStreamSynchronize(s));
    if (h_err) return fail("h2_poly_lookup_permute: an input value does not occur in the table (Error::ConstraintSystemFailure, plonk/lookup/prover.rs:605-608)");
    return 0;
}
extern "C" int h2_poly_lookup_permute(uint64_t input, uint64_t table, size_t usable_rows, uint64_t out_input, uint64_t out_table) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    PolyBuf *a = find_poly(input), *t = find_poly(table), *oa = find_poly(out_input), *ot = find_poly(out_table);
    if (!a || !t || !oa || !ot) return fail("h2_poly_lookup_permute: unknown polynomial handle");
    if (oa == ot || oa == a || oa == t || ot == a || ot == t) return fail("h2_poly_lookup_permute: the outputs must be two polynomials other than the inputs");
    if (a->field != t->field || a->field != oa->field || a->field != ot->field) return fail("h2_poly_lookup_permute: the polynomials live in different fields");
    if (a->len < usable_rows || t->len < usable_rows || oa->len < usable_rows || ot->len < usable_rows)
        return fail("h2_poly_lookup_permute: a polynomial holds fewer than usable_rows elements");
    if (usable_rows >= (1ull << 31)) return fail("h2_poly_lookup_permute: usable_rows >= 2^31");
    if (usable_rows == 0) return 0;
    if (a->field == H2_FIELD_FP) return lookup_permute_run<FpParams>(a, t, usable_rows, oa, ot);
    return lookup_permute_run<FqParams>(a, t, usable_rows, oa, ot);
}


// ------------------------------------------------------------------------------------------------
// the verifier's MSM: g_scalars resident (verifier.cuh)
// ------------------------------------------------------------------------------------------------
// dst[i] (+)= init * prod_{j : bit j of i} u[k - 1 - j], i < 2^k: compute_s (poly/commitment/verifier.rs:156-171); with
// `accumulate` the add_to_g_scalars of Guard::use_challenges (:36-41, msm.rs:104-113) in the same pass
template <class P> static int compute_s_run(PolyBuf *d, const void *u, uint32_t k, const void *init, int accumulate, int repr) {
    Context &X = g_ctx;
    cudaStream_t s = X.stream;
    if (scratch_acquire(s)) return 1;
    if (X.po_pts.ensure((size_t)k * sizeof(fe))) return 1;
    fe *du = X.po_pts.as<fe>();
    CU(cudaMemcpyAsync(du, u, (size_t)k * sizeof(fe), cudaMemcpyHostToDevice, s));
    if (repr == H2_REPR_CANONICAL) LAUNCH(convert_kernel<P>, blocks_for(k, 64), 64, 0, s, du, (uint64_t)k, 1);
    const uint64_t groups = 1ull << (k - (k < 2 ? k : 2));
    LAUNCH(verifier_compute_s_kernel<P>, blocks_for(groups, 128), 128, 0, s, d->buf.as<fe>(), (const fe *)du, k, host_to_mont<P>(init, repr), accumulate);
    return scratch_release(s);
}
extern "C" int h2_poly_compute_s(uint64_t dst, const void *u, uint32_t k, const void *init, int accumulate, int repr) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    PolyBuf *d = find_poly(dst);
    if (!d) return fail("h2_poly_compute_s: unknown polynomial handle");
    if (!u || !init) return fail("h2_poly_compute_s: null challenge vector or init");
    if (k == 0) return fail("h2_poly_compute_s: no challenges (assert!(!u.is_empty()), poly/commitment/verifier.rs:157)");
    if (k > 30 || d->len < ((size_t)1 << k)) return fail("h2_poly_compute_s: the polynomial holds fewer than 2^k elements");
    if (d->field == H2_FIELD_FP) return compute_s_run<FpParams>(d, u, k, init, accumulate, repr);
    return compute_s_run<FqParams>(d, u, k, init, accumulate, repr);
}
// dst[i] = a * dst[i] + b * src[i], i < n (src == 0: dst[i] *= a): MSM::scale and the g_scalars part of MSM::add_msm
// (poly/commitment/msm.rs:126-139, :37-62); BatchVerifier's accumulate_msm (plonk/verifier/batch.rs:83-93) is one call
extern "C" int h2_poly_scale_add(uint64_t dst, const void *a, uint64_t src, const void *b, size_t n, int repr) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (require_ready()) return 1;
    PolyBuf *d = find_poly(dst), *x = src ? find_poly(src) : nullptr;
    if (!d || (src && !x)) return fail("h2_poly_scale_add: unknown polynomial handle");
    if (x == d) return fail("h2_poly_scale_add: src must be another polynomial than dst");
    if (x && x->field != d->field) return fail("h2_poly_scale_add: the polynomials live in different fields");
    if (d->len < n || (x && x->len < n)) return fail("h2_poly_scale_add: a polynomial holds fewer than n elements");
    if (!a || (x && !b)) return fail("h2_poly_scale_add: null factor");
    if (n == 0) return 0;
    cudaStream_t s = g_ctx.stream;
    const fe *sp = x ? x->buf.as<fe>() : nullptr;
    if (d->field == H2_FIELD_FP)
        LAUNCH(verifier_scale_add_kernel<FpParams>, blocks_for(n, 256), 256, 0, s, d->buf.as<fe>(), sp, host_to_mont<FpParams>(a, repr),
               x ? host_to_mont<FpParams>(b, repr) : fe_zero(), (uint64_t)n);
    else
        LAUNCH(verifier_scale_add_kernel<FqParams>, blocks_for(n, 256), 256, 0, s, d->buf.as<fe>(), sp, host_to_mont<FqParams>(a, repr),
               x ? host_to_mont<FqParams>(b, repr) : fe_zero(), (uint64_t)n);
    return 0;
}
