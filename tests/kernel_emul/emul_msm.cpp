// TEST-ONLY serial execution of the MSM kernel bodies (msm.cuh) on the host: checks the
// digit/sort/work-item/partial/reduce/combine index logic against the oracle without a GPU.
#include <cstring>
#include <vector>
#include "msm.cuh"
using namespace h2;

static uint32_t g_force_cap = 0;      // 0 auto, H2_MSM_NO_BINS exact sort only, else the bin capacity (tests force overflows)
extern "C" void emu_msm_set_cap(uint32_t cap) { g_force_cap = cap; }
static uint32_t g_ba_rounds = 0, g_ba_target = 64;   // batched-affine rounds before the XYZZ chain (0: classic accumulation)
extern "C" void emu_msm_set_ba(uint32_t rounds, uint32_t target) { g_ba_rounds = rounds; g_ba_target = target ? target : 64; }
static uint32_t g_last_ba = 0;
extern "C" uint32_t emu_msm_last_ba(void) { return g_last_ba; }
template <class P, class PS>
static int run_msm(const uint8_t *scalars, const uint8_t *bases, size_t n, uint32_t c, int scalars_mont,
                   uint32_t force_t, uint32_t force_kn, int fixed, int glv, uint8_t *out_xyz, uint32_t sets = 1) {
    // batched fixed-base mode: `scalars` holds sets * n scalars, `out_xyz` receives sets * 96 bytes
    const size_t ns = n * sets;
    MsmPlan p;
    if (!c) c = msm_default_window(n, glv ? 1u : 0u);
    if (fixed && c < 4) c = 4;      // table windows: W = ceil(256 / c) <= 64
    msm_make_plan(p, n, c, force_t, force_kn, fixed ? 1u : 0u, n + 3, glv ? 1u : 0u, sets, g_force_cap);
    if (p.acc_levels > H2_MSM_MAX_LEVELS) return -2;
    msm_plan_ba(p, g_ba_rounds, g_ba_target);
    g_last_ba = p.ba;
    std::vector<fe> sc(ns ? ns : 1), sc_canon(ns ? ns : 1);
    std::vector<affine> bs(n ? n : 1);
    for (size_t i = 0; i < ns; i++) {
        memcpy(sc[i].v, scalars + 32 * i, 32);
        if (scalars_mont) sc[i] = fe_to_mont<PS>(sc[i]);
    }
    for (size_t i = 0; i < n; i++) {
        affine a; memcpy(a.x.v, bases + 64 * i, 32); memcpy(a.y.v, bases + 64 * i + 32, 32);
        if (!affine_is_identity(a)) { a.x = fe_to_mont<P>(a.x); a.y = fe_to_mont<P>(a.y); }
        bs[i] = a;
    }
    std::vector<affine> phi(n ? n : 1);
    for (size_t i = 0; i < n; i++) { phi[i] = bs[i]; phi[i].x = fe_mul<P>(bs[i].x, glv_zeta<P>()); }
    std::vector<affine> table;
    if (fixed) {   // table[w * stride + i] = 2^(c w) * base[i]; stride > n on purpose
        table.resize((size_t)p.W * p.stride);
        for (size_t i = 0; i < n; i++) Msm<P, PS>::table_body(bs.data(), table.data(), n, p.stride, p.c, p.W, i);
    }
    std::vector<uint32_t> counts(p.G + 1, 0), cursor(p.G, 0), cursor2(p.G, 0), refs(p.ref_space ? p.ref_space : 1), size_hist(p.T + 2, 0),
        size_cursor(p.T + 1, 0), flags(8, 0);
    std::vector<uint2> items(p.max_items);
    std::vector<xyzz> bucket_sum(p.G, xyzz_identity());
    size_t pt = p.part_total ? p.part_total : 1;
    std::vector<uint32_t> pkey(pt, H2_MSM_INVALID_KEY), pstart(pt), pend(pt);
    std::vector<xyzz> ppt(pt), ra_t((size_t)p.W * p.m1), ra_e((size_t)p.W * p.m1), r0((size_t)p.W * p.nb0 * H2_R0_ROWS),
        r1((size_t)p.W * p.r1_rows), wsum(p.W);
    std::vector<jacobian> result(sets);
    MsmBuffers M;
    std::vector<uint32_t> glv_parts((ns ? ns : 1) * 8 + 4);
    M.glv_parts = glv_parts.data();
    M.scalars = sc.data(); M.bases = fixed ? table.data() : bs.data(); M.bases_phi = phi.data(); M.scalars_mont = scalars_mont; M.scal_canon = sc_canon.data();
    M.counts = counts.data(); M.cursor = cursor.data(); M.cursor2 = cursor2.data(); M.refs = refs.data();
    M.size_hist = size_hist.data(); M.size_cursor = size_cursor.data(); M.flags = flags.data(); M.items = items.data();
    M.bucket_sum = bucket_sum.data(); M.pkey = pkey.data(); M.pstart = pstart.data(); M.pend = pend.data();
    M.ppt = ppt.data(); M.ra_t = ra_t.data(); M.ra_e = ra_e.data(); M.r0 = r0.data(); M.r1 = r1.data();
    M.wsum = wsum.data(); M.result = result.data();
    std::vector<affine> ba_lv[H2_BA_MAX_ROUNDS];
    for (uint32_t r = 0; r < H2_BA_MAX_ROUNDS; r++) {
        affine poison; for (int i = 0; i < 8; i++) { poison.x.v[i] = 0xdeadbeefu; poison.y.v[i] = 0xdeadbeefu; }
        ba_lv[r].assign(p.ba ? (p.ref_space >> (r + 1)) + 1 : 1, poison);
        M.ba[r] = ba_lv[r].data();
    }
    typedef Msm<P, PS> K;
    // single-pass binned sort (reverse order to mimic the arbitrary order atomics give); a full bin -> flags[1]
    if (p.cap == 0) flags[1] = 1;
    else
        for (size_t ii = ns; ii-- > 0;) {
            bool ok = K::for_each_digit(p, M, ii, true, [&](uint32_t g, uint32_t ref) {
                uint32_t slot = cursor[g]++, lo, cap;
                if (K::bin_of(p, g, lo, cap) && slot < cap) refs[lo + slot] = ref; else flags[1] = 1;
            });
            if (!ok) return -3;
        }
    if (flags[1]) {   // exact sort: K2 histogram, scan, K3 scatter
        for (size_t i = 0; i < ns; i++) {
            bool ok = K::for_each_digit(p, M, i, true, [&](uint32_t g, uint32_t) { counts[g]++; });
            if (!ok) return -3;
        }
        uint32_t run = 0;
        for (uint64_t g = 0; g <= p.G; g++) { uint32_t v = counts[g]; counts[g] = run; run += v; }
        for (size_t ii = ns; ii-- > 0;) {
            K::for_each_digit(p, M, ii, false, [&](uint32_t g, uint32_t ref) { refs[counts[g] + cursor2[g]++] = ref; });
        }
    }
    // work items
    for (uint64_t g = 0; g < p.G; g++) {
        uint32_t nfull, rem; K::count_items(p, M, g, nfull, rem);
        size_hist[p.T] += nfull; if (rem) size_hist[rem]++;
    }
    K::size_bases_body(p, M);
    for (uint64_t g = p.G; g-- > 0;) {
        uint32_t nfull, rem; K::count_items(p, M, g, nfull, rem);
        for (uint32_t k = 0; k < nfull; k++) items[size_cursor[p.T]++] = make_uint2((uint32_t)g, K::bucket_lo(p, M, g) + k * p.T);
        if (rem) items[size_cursor[rem]++] = make_uint2((uint32_t)g, K::bucket_lo(p, M, g) + nfull * p.T);
    }
    if (size_hist[p.T + 1] > p.max_items) return -4;
    // K4a: batched-affine rounds + the chain over the last level (no-ops when the exact sort ran)
    for (uint32_t r = 1; r <= p.ba; r++)
        for (uint64_t j = 0; j < (p.max_items + p.ba_m[r - 1] - 1) / p.ba_m[r - 1]; j++) K::template ba_round_body<8, 4>(p, M, r, j);
    if (p.ba) for (uint64_t t = 0; t < p.max_items; t++) K::accum0_pts_body(p, M, t);
    // K4
    for (uint64_t t = 0; t < p.max_items; t++) K::accum0_body(p, M, t);
    for (uint32_t lv = 1; lv < p.acc_levels; lv++)
        for (uint64_t t = 0; t < p.acc_threads[lv]; t++) K::accumN_body(p, M, lv, t);
    // K5
    for (uint64_t t = 0; t < (uint64_t)p.Wb * p.m1; t++) K::reduceA_body(p, M, t);
    for (uint32_t w = 0; w < p.Wb; w++)
        for (uint32_t blk = 0; blk < p.nb0; blk++)
            for (uint32_t row = 0; row < H2_R0_ROWS; row++) {
                xyzz v = xyzz_identity();
                if (row < 2 + p.bits0)
                    for (uint32_t lane = 0; lane < (1u << H2_R0_LOG); lane++) { xyzz cc = K::r0_contrib(p, M, w, blk, row, lane); xyzz_add<P>(v, cc); }
                r0[((size_t)w * p.nb0 + blk) * H2_R0_ROWS + row] = v;
            }
    for (uint32_t w = 0; w < p.Wb; w++)
        for (uint32_t row = 0; row < p.r1_rows; row++) {
            xyzz v = xyzz_identity();
            for (uint32_t blk = 0; blk < p.nb0; blk++) { xyzz cc = K::r1_contrib(p, M, w, row, blk); xyzz_add<P>(v, cc); }
            r1[(size_t)w * p.r1_rows + row] = v;
        }
    xyzz total = xyzz_identity();
    for (uint32_t w = 0; w < p.Wb; w++) {
        xyzz ws = xyzz_identity();
        for (uint32_t r = 0; r < 32; r++) { xyzz cc = K::wsum_item(p, M, w, r); xyzz_add<P>(ws, cc); }
        if (p.fixed) K::finish(M, ws, 1, w); else xyzz_add<P>(total, ws);
    }
    if (!p.fixed) K::finish(M, total, 1);
    for (uint32_t k = 0; k < sets; k++) {
        memcpy(out_xyz + 96 * k, result[k].x.v, 32); memcpy(out_xyz + 96 * k + 32, result[k].y.v, 32); memcpy(out_xyz + 96 * k + 64, result[k].z.v, 32);
    }
    return (int)p.acc_levels + (flags[0] ? 100 : 0) + (flags[1] ? 1000 : 0);
}

// curve 0 = Pallas (coords Fp, scalars Fq), 1 = Vesta.  Returns acc_levels (+100 if some bucket was split, +1000 if the
// exact sort ran) or <0.
extern "C" int emu_msm(int curve, const uint8_t *scalars, const uint8_t *bases, size_t n, uint32_t c,
                       int scalars_mont, uint32_t force_t, uint32_t force_kn, uint8_t *out_xyz) {
    if (curve == 0) return run_msm<FpParams, FqParams>(scalars, bases, n, c, scalars_mont, force_t, force_kn, 0, 0, out_xyz);
    return run_msm<FqParams, FpParams>(scalars, bases, n, c, scalars_mont, force_t, force_kn, 0, 0, out_xyz);
}
// same MSM with the GLV endomorphism split
extern "C" int emu_msm_glv(int curve, const uint8_t *scalars, const uint8_t *bases, size_t n, uint32_t c,
                           int scalars_mont, uint32_t force_t, uint32_t force_kn, uint8_t *out_xyz) {
    if (curve == 0) return run_msm<FpParams, FqParams>(scalars, bases, n, c, scalars_mont, force_t, force_kn, 0, 1, out_xyz);
    return run_msm<FqParams, FpParams>(scalars, bases, n, c, scalars_mont, force_t, force_kn, 0, 1, out_xyz);
}
// same MSM through the precomputed window table (resident-bases path of Params::commit*)
extern "C" int emu_msm_fixed(int curve, const uint8_t *scalars, const uint8_t *bases, size_t n, uint32_t c,
                             uint32_t force_t, uint32_t force_kn, uint8_t *out_xyz) {
    if (curve == 0) return run_msm<FpParams, FqParams>(scalars, bases, n, c, 0, force_t, force_kn, 1, 0, out_xyz);
    return run_msm<FqParams, FpParams>(scalars, bases, n, c, 0, force_t, force_kn, 1, 0, out_xyz);
}
// batched: `sets` scalar vectors of n entries against the same table
extern "C" int emu_msm_fixed_batch(int curve, const uint8_t *scalars, const uint8_t *bases, size_t n, uint32_t sets, uint32_t c,
                                   uint8_t *out_xyz) {
    if (curve == 0) return run_msm<FpParams, FqParams>(scalars, bases, n, c, 0, 0, 0, 1, 0, out_xyz, sets);
    return run_msm<FqParams, FpParams>(scalars, bases, n, c, 0, 0, 0, 1, 0, out_xyz, sets);
}

// ---- IPA round loop (ipa.cuh bodies + the 2-set fixed-base MSM above) ------------------------------
#include "ipa.cuh"
template <class P, class PS>
static int run_ipa(const uint8_t *bases, uint32_t k, const uint8_t *p_prime, const uint8_t *x3, const uint8_t *z, const uint8_t *chal,
                   const uint8_t *chal_inv, const uint8_t *l_rand, const uint8_t *r_rand, uint32_t c, uint8_t *out_l, uint8_t *out_r, uint8_t *out_c) {
    const uint64_t n = 1ull << k;
    auto rd = [](const uint8_t *b) { fe x; memcpy(x.v, b, 32); return fe_to_mont<PS>(x); };
    std::vector<fe> p(n), b(n), s(n), scal(2 * (n + 2));
    for (uint64_t i = 0; i < n; i++) memcpy(p[i].v, p_prime + 32 * i, 32);
    IpaState S; S.p = p.data(); S.b = b.data(); S.s = s.data(); S.scal = scal.data(); S.n = n;
    for (uint64_t t = 0; t < n; t++) Ipa<PS>::init_body(S, 0, t);
    fe x = rd(x3), cur = fe_one<PS>();
    for (uint64_t t = 0; t < n; t++) { b[t] = cur; cur = fe_mul<PS>(cur, x); }
    fe zz = rd(z);
    std::vector<uint8_t> sc_bytes(2 * (n + 2) * 32), lr(2 * 96);
    for (uint32_t j = 0; j < k; j++) {
        const uint32_t bit = k - 1 - j;
        for (uint64_t t = 0; t < n; t++) Ipa<PS>::prep_body(S, bit, t);
        fe vl = fe_zero(), vr = fe_zero();
        const uint32_t nthr = 7;
        for (uint32_t tid = 0; tid < nthr; tid++) { fe a, bb; Ipa<PS>::inner_partial(S, bit, tid, nthr, a, bb); vl = fe_add<PS>(vl, a); vr = fe_add<PS>(vr, bb); }
        Ipa<PS>::inner_finish(S, vl, vr, zz, rd(l_rand + 32 * j), rd(r_rand + 32 * j));
        for (uint64_t i = 0; i < 2 * (n + 2); i++) { fe v = fe_from_mont<PS>(scal[i]); memcpy(&sc_bytes[32 * i], v.v, 32); }
        int rc = run_msm<P, PS>(sc_bytes.data(), bases, n + 2, c, 0, 0, 0, 1, 0, lr.data(), 2);
        if (rc < 0) return rc;
        memcpy(out_l + 96 * j, lr.data(), 96); memcpy(out_r + 96 * j, lr.data() + 96, 96);
        fe u = rd(chal + 32 * j), ui = rd(chal_inv + 32 * j);
        for (uint64_t t = 0; t < n; t++) Ipa<PS>::fold_body(S, bit, u, ui, t);
    }
    fe cc = fe_from_mont<PS>(p[0]);
    memcpy(out_c, cc.v, 32);
    return 0;
}
extern "C" int emu_ipa(int curve, const uint8_t *bases, uint32_t k, const uint8_t *p_prime, const uint8_t *x3, const uint8_t *z, const uint8_t *chal,
                       const uint8_t *chal_inv, const uint8_t *l_rand, const uint8_t *r_rand, uint32_t c, uint8_t *out_l, uint8_t *out_r, uint8_t *out_c) {
    if (curve == 0) return run_ipa<FpParams, FqParams>(bases, k, p_prime, x3, z, chal, chal_inv, l_rand, r_rand, c, out_l, out_r, out_c);
    return run_ipa<FqParams, FpParams>(bases, k, p_prime, x3, z, chal, chal_inv, l_rand, r_rand, c, out_l, out_r, out_c);
}

// plan introspection for the tests: out = {W, B, G, T, cap, l0, cap_top, top_bins}
extern "C" void emu_msm_plan(size_t n, uint32_t c, int fixed, int glv, uint32_t sets, uint32_t force_cap, uint64_t *out) {
    MsmPlan p;
    msm_make_plan(p, n, c, 0, 0, fixed ? 1u : 0u, n + 3, glv ? 1u : 0u, sets, force_cap);
    out[0] = p.W; out[1] = p.B; out[2] = p.G; out[3] = p.T; out[4] = p.cap; out[5] = p.l0; out[6] = p.cap_top; out[7] = p.top_bins;
}
