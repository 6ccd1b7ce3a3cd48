// TEST-ONLY serial execution of the MSM kernel bodies (msm.cuh) on the host: checks the
// digit/sort/chunk/partial/reduce index logic against the oracle without a GPU.
#include <cstring>
#include <vector>
#include "msm.cuh"
using namespace h2;

template <class P, class PS>
static int run_msm(const uint8_t *scalars, const uint8_t *bases, size_t n, uint32_t c, int scalars_mont,
                   uint32_t force_k0, uint8_t *out_xyz) {
    MsmPlan p;
    msm_make_plan(p, n, c ? c : msm_default_window(n));
    if (force_k0) {   // re-plan with a forced level-0 chunk to exercise multi-level partial merging
        p.acc_chunk[0] = force_k0;
        uint32_t lv = 0; uint64_t slots = p.max_refs; p.part_total = 0;
        for (;;) {
            uint32_t chunk = lv == 0 ? force_k0 : 8u;
            uint64_t threads = (slots + chunk - 1) / chunk; if (!threads) threads = 1;
            p.acc_chunk[lv] = chunk; p.acc_threads[lv] = threads; p.acc_slots[lv] = slots; lv++;
            if (threads == 1) break;
            p.part_offset[lv] = p.part_total; slots = 2 * threads; p.part_total += slots;
            if (lv >= H2_MSM_MAX_LEVELS) return -2;
        }
        p.acc_levels = lv;
    }
    std::vector<fe> sc(n ? n : 1), sc_canon(n ? n : 1);
    std::vector<affine> bs(n ? n : 1);
    for (size_t i = 0; i < n; i++) {
        memcpy(sc[i].v, scalars + 32 * i, 32);
        if (scalars_mont) sc[i] = fe_to_mont<PS>(sc[i]);
        affine a; memcpy(a.x.v, bases + 64 * i, 32); memcpy(a.y.v, bases + 64 * i + 32, 32);
        if (!affine_is_identity(a)) { a.x = fe_to_mont<P>(a.x); a.y = fe_to_mont<P>(a.y); }
        bs[i] = a;
    }
    std::vector<uint32_t> counts(p.G + 1, 0), cursor(p.G, 0), refs(p.max_refs ? p.max_refs : 1), keys(p.max_refs ? p.max_refs : 1);
    std::vector<xyzz> bucket_sum(p.G, xyzz_identity());
    size_t pt = p.part_total ? p.part_total : 1;
    std::vector<uint32_t> pkey(pt, H2_MSM_INVALID_KEY), pstart(pt), pend(pt);
    std::vector<xyzz> ppt(pt), red_sums(p.red_total), red_e(p.red_total);
    jacobian result;
    MsmBuffers M;
    M.scalars = sc.data(); M.bases = bs.data(); M.scalars_mont = scalars_mont; M.scal_canon = sc_canon.data();
    M.counts = counts.data(); M.cursor = cursor.data(); M.refs = refs.data(); M.keys = keys.data();
    M.bucket_sum = bucket_sum.data(); M.pkey = pkey.data(); M.pstart = pstart.data(); M.pend = pend.data();
    M.ppt = ppt.data(); M.red_sums = red_sums.data(); M.red_e = red_e.data(); M.win_sums = nullptr; M.result = &result;
    typedef Msm<P, PS> K;
    // K2 histogram
    for (size_t i = 0; i < n; i++) {
        uint32_t s[8]; K::load_scalar(M, i, s, true);
        uint32_t carry = 0;
        for (uint32_t w = 0; w < p.W; w++) {
            int32_t d = K::next_digit(s, w, p.c, carry);
            if (d) counts[(uint64_t)w * p.B + (uint32_t)(d < 0 ? -d : d) - 1]++;
        }
        if (carry) return -3;    // top window must absorb the carry
    }
    // scan
    uint32_t run = 0;
    for (uint64_t g = 0; g <= p.G; g++) { uint32_t v = counts[g]; counts[g] = run; run += v; }
    // K3 scatter (reverse order to mimic the arbitrary order atomics give)
    for (size_t ii = n; ii-- > 0;) {
        uint32_t s[8]; K::load_scalar(M, ii, s, false);
        uint32_t carry = 0;
        for (uint32_t w = 0; w < p.W; w++) {
            int32_t d = K::next_digit(s, w, p.c, carry);
            if (!d) continue;
            uint64_t g = (uint64_t)w * p.B + (uint32_t)(d < 0 ? -d : d) - 1;
            uint32_t pos = counts[g] + cursor[g]++;
            refs[pos] = (uint32_t)ii | (d < 0 ? 0x80000000u : 0u);
            keys[pos] = (uint32_t)g;
        }
    }
    // K4 accumulate levels
    for (uint64_t t = 0; t < p.acc_threads[0]; t++) K::accum0_body(p, M, t);
    for (uint32_t lv = 1; lv < p.acc_levels; lv++)
        for (uint64_t t = 0; t < p.acc_threads[lv]; t++) K::accumN_body(p, M, lv, t);
    // K5 reduce + combine
    for (uint32_t lv = 0; lv < p.red_levels; lv++) {
        uint32_t m_out = (p.red_m_in[lv] + (1u << p.red_log_l[lv]) - 1) >> p.red_log_l[lv];
        for (uint64_t t = 0; t < (uint64_t)p.W * m_out; t++) K::reduce_body(p, M, lv, t);
    }
    xyzz total = xyzz_identity();
    for (uint32_t w = 0; w < p.W; w++) { xyzz v = K::window_value(p, M, w); xyzz_add<P>(total, v); }
    K::finish(M, total, 1);
    memcpy(out_xyz, result.x.v, 32); memcpy(out_xyz + 32, result.y.v, 32); memcpy(out_xyz + 64, result.z.v, 32);
    return (int)p.acc_levels;
}

// curve 0 = Pallas (coords Fp, scalars Fq), 1 = Vesta.  Returns the number of accumulate levels (>0) or <0.
extern "C" int emu_msm(int curve, const uint8_t *scalars, const uint8_t *bases, size_t n, uint32_t c,
                       int scalars_mont, uint32_t force_k0, uint8_t *out_xyz) {
    if (curve == 0) return run_msm<FpParams, FqParams>(scalars, bases, n, c, scalars_mont, force_k0, out_xyz);
    return run_msm<FqParams, FpParams>(scalars, bases, n, c, scalars_mont, force_k0, out_xyz);
}
