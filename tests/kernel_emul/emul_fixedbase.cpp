// TEST-ONLY serial execution of the direct-sum fixed-base MSM (fixedbase.cuh) on the host: window table (c = 8),
// multiples table, accumulate threads, reduce levels exactly as capi.cu schedules them.
#include <cstring>
#include <vector>
#include "fixedbase.cuh"
using namespace h2;

// scalars: [sets][total] canonical (or Montgomery when scalars_mont); bases: n >= total canonical affine; out: sets x 64 B affine
template <class P, class PS>
static int run_fb(const uint8_t *scalars, const uint8_t *bases, size_t n, size_t total, uint32_t sets, uint32_t split, int scalars_mont, uint8_t *out) {
    std::vector<affine> bs(n), wtab((size_t)H2_FB_WINDOWS * n), dtab((size_t)H2_FB_WINDOWS * H2_FB_MULTIPLES * n);
    for (size_t i = 0; i < n; i++) {
        memcpy(&bs[i], bases + 64 * i, 64);
        if (!affine_is_identity(bs[i])) { bs[i].x = fe_to_mont<P>(bs[i].x); bs[i].y = fe_to_mont<P>(bs[i].y); }
    }
    for (size_t i = 0; i < n; i++) Msm<P, PS>::table_body(bs.data(), wtab.data(), n, n, H2_FB_BITS, H2_FB_WINDOWS, i);
    for (uint64_t t = 0; t < (uint64_t)H2_FB_WINDOWS * n; t++) FixedBase<P, PS>::table_body(wtab.data(), dtab.data(), n, n, t);
    std::vector<fe> sc(sets * total);
    memcpy(sc.data(), scalars, sets * total * 32);
    FbPlan p;
    p.total = total; p.sets = sets; p.split = split ? split : fb_split(total, sets); p.scalars_mont = scalars_mont ? 1u : 0u;
    uint64_t count = total * p.split;
    std::vector<xyzz> a(sets * count), b;
    for (uint64_t u = 0; u < sets * count; u++) FixedBase<P, PS>::accum_body(p, sc.data(), dtab.data(), a.data(), u);
    uint64_t in_stride = count;
    for (;;) {
        const uint32_t f = fb_fan(count);
        const uint64_t ctas = fb_ctas(count, f);
        b.assign(sets * ctas, xyzz_identity());
        for (uint32_t set = 0; set < sets; set++)
            for (uint64_t cta = 0; cta < ctas; cta++) {
                xyzz sm[H2_FB_QUADS];
                for (uint32_t q = 0; q < H2_FB_QUADS; q++) sm[q] = FixedBase<P, PS>::reduce_gather(a.data() + set * in_stride, count, f, cta, q);
                for (uint32_t step = H2_FB_QUADS / 2; step >= 1; step >>= 1)
                    for (uint32_t q = 0; q < step; q++) sm[q] = FixedBase<P, PS>::reduce_level(sm[q], sm[q + step]);
                b[set * ctas + cta] = sm[0];
            }
        a.swap(b);
        count = ctas; in_stride = ctas;
        if (ctas == 1) break;
    }
    for (uint32_t set = 0; set < sets; set++) {
        jacobian j;
        FixedBase<P, PS>::finish(&j, a[set], 0);
        affine r = jacobian_to_affine<P>(j);
        r.x = fe_from_mont<P>(r.x); r.y = fe_from_mont<P>(r.y);
        memcpy(out + 64 * set, &r, 64);
    }
    return (int)p.split;
}
extern "C" int emu_msm_direct(int curve, const uint8_t *scalars, const uint8_t *bases, size_t n, size_t total, uint32_t sets, uint32_t split,
                              int scalars_mont, uint8_t *out) {
    if (curve == 0) return run_fb<FpParams, FqParams>(scalars, bases, n, total, sets, split, scalars_mont, out);
    return run_fb<FqParams, FpParams>(scalars, bases, n, total, sets, split, scalars_mont, out);
}
