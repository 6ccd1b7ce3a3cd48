// GLV endomorphism split for the j = 0 Pasta curves (generated constants: tools/gen_glv_constants.py).
//
// phi(x, y) = (zeta x, y) equals multiplication by lambda, so  k P = k1 P + k2 phi(P)  with
// k = k1 + k2 lambda (mod r) and |k1|, |k2| < 2^127.  A 255-bit MSM over n points becomes a 127-bit
// MSM over 2n points: the same number of (point, window) references, but half the windows --
// half the bucket sets to reduce and half the serial doubling chain of the window combine
// (arithmetic.rs:163 does c*i doublings per window; that chain is the latency floor of an MSM).
// The group element computed is the same as best_multiexp's.
//
// k1 = k - c1 a1 - c2 a2,  k2 = c1 |b1| - c2 b2  with (a1, b1), (a2, b2) a reduced lattice basis
// (b1 < 0 < a1, a2, b2) and c_i = floor((k g_i + 2^383) / 2^384), g1 = round(2^384 b2 / r), g2 = round(2^384 |b1| / r):
// c_i differs from the real k b / r by at most 1/2 + eps, eps < 2^-130 (|g_i - 2^384 b / r| <= 1/2 and k < 2^255), so
// |k1| <= (1/2 + eps)(a1 + a2) and |k2| <= (1/2 + eps)(|b1| + b2); for both curves these sums are 0.577 / 0.866 x 2^128,
// hence |k_i| < 0.867 x 2^127 < 2^127 ALWAYS: four limbs per half with a free top bit (it carries the sign in the stored
// form), and W = ceil(128 / c) windows (tests/test_kernel_emul.py::test_emul_glv_split_bounds).  The identity
// k1 + k2 lambda = k holds for any integers c1, c2; the rounding only bounds the size.
#pragma once
#include "field.cuh"

namespace h2 {

template <class P> struct GlvConst;   // keyed by the curve's COORDINATE field (FpParams = Pallas)

// pallas: lambda = 0x6819a58283e528e511db4d81cf70f5a0fed467d47c033af2aa9d2e050aa0e4f
//          zeta = 0x12ccca834acdba712caad5dc57aab1b01d1f8bd237ad31491dad5ebdfdfe4ab9   (max |k_i| bits observed: 127)
template <> struct GlvConst<FpParams> {
    static H2_HD uint32_t zeta_mont(int i) { constexpr uint32_t v[8] = {0x619a153du, 0x02021cf6u, 0x4980b78eu, 0x9e8c2697u, 0xc87a4666u, 0x2a676d5cu, 0xa7a17876u, 0x15d8049du}; return v[i]; }
    static H2_HD uint32_t a1(int i) { constexpr uint32_t v[4] = {0x00000001u, 0x7fcae1c7u, 0x40f04915u, 0x49e69d16u}; return v[i]; }
    static H2_HD uint32_t nb1(int i) { constexpr uint32_t v[4] = {0x00000000u, 0x8cb12793u, 0x40a89953u, 0x49e69d16u}; return v[i]; }
    static H2_HD uint32_t a2(int i) { constexpr uint32_t v[4] = {0x00000000u, 0x8cb12793u, 0x40a89953u, 0x49e69d16u}; return v[i]; }
    static H2_HD uint32_t b2(int i) { constexpr uint32_t v[4] = {0x00000001u, 0x0c7c095au, 0x8198e269u, 0x93cd3a2cu}; return v[i]; }
    static H2_HD uint32_t g1(int i) { constexpr uint32_t v[9] = {0x11afc293u, 0x111f6861u, 0x086862e0u, 0xc35fbd4du, 0x00000002u, 0x31f02568u, 0x066389a4u, 0x4f34e8b2u, 0x00000002u}; return v[i]; }
    static H2_HD uint32_t g2(int i) { constexpr uint32_t v[9] = {0x72171db4u, 0x4a95a2d9u, 0x8480fa55u, 0x61afdea6u, 0xffffffffu, 0x32c49e4bu, 0x02a2654eu, 0x279a7459u, 0x00000001u}; return v[i]; }
};
// vesta: lambda = 0x2d33357cb532458ed3552a23a8554e5005270d29d19fc7d27b7fd22f0201b547
//          zeta = 0x397e65a7d7c1ad71aee24b27e308f0a61259527ec1d4752e619d1840af55f1b1   (max |k_i| bits observed: 127)
template <> struct GlvConst<FqParams> {
    static H2_HD uint32_t zeta_mont(int i) { constexpr uint32_t v[8] = {0x7feeeee3u, 0x410e7d20u, 0xd8fa2279u, 0x6afdf14fu, 0xeca4d4d7u, 0xfd3d8a04u, 0x77dba4efu, 0x2de2d607u}; return v[i]; }
    static H2_HD uint32_t a1(int i) { constexpr uint32_t v[4] = {0x00000001u, 0x8cb12793u, 0x40a89953u, 0x49e69d16u}; return v[i]; }
    static H2_HD uint32_t nb1(int i) { constexpr uint32_t v[4] = {0x00000000u, 0x7fcae1c7u, 0x40f04915u, 0x49e69d16u}; return v[i]; }
    static H2_HD uint32_t a2(int i) { constexpr uint32_t v[4] = {0x00000001u, 0x0c7c095au, 0x8198e269u, 0x93cd3a2cu}; return v[i]; }
    static H2_HD uint32_t b2(int i) { constexpr uint32_t v[4] = {0x00000001u, 0x8cb12793u, 0x40a89953u, 0x49e69d16u}; return v[i]; }
    static H2_HD uint32_t g1(int i) { constexpr uint32_t v[9] = {0x4bf99a83u, 0x841414c2u, 0x85cc1578u, 0x61afdea6u, 0x00000003u, 0x32c49e4cu, 0x02a2654eu, 0x279a7459u, 0x00000001u}; return v[i]; }
    static H2_HD uint32_t g2(int i) { constexpr uint32_t v[9] = {0xdd747ae0u, 0x0009789fu, 0x853283aeu, 0x61afdea6u, 0xffffffffu, 0xff2b871bu, 0x03c12455u, 0x279a7459u, 0x00000001u}; return v[i]; }
};

namespace glv {
// out[0..NO) = limbs [LO, LO + NO) of a (NA limbs) * b (NB limbs)
template <int NA, int NB, int LO, int NO> H2_HD void mul_part(const uint32_t *a, const uint32_t *b, uint32_t *out) {
    uint64_t lo = 0;      // running column sum: low 64 bits
    uint32_t hi = 0;      // and its overflow
    for (int col = 0; col < LO + NO; col++) {
        for (int i = 0; i < NA; i++) {
            int j = col - i;
            if (j < 0 || j >= NB) continue;
            uint64_t p = (uint64_t)a[i] * b[j];
            lo += p;
            if (lo < p) hi++;
        }
        if (col >= LO) out[col - LO] = (uint32_t)lo;
        lo = (lo >> 32) | ((uint64_t)hi << 32);
        hi = 0;
    }
}
// c = (w + 2^31) >> 32 on a 6-limb value
H2_HD void round_shift(const uint32_t *w, uint32_t *c) {
    uint64_t carry = ((uint64_t)w[0] + 0x80000000u) >> 32;
    for (int i = 0; i < 5; i++) { carry += w[i + 1]; c[i] = (uint32_t)carry; carry >>= 32; }
}
H2_HD void add8(uint32_t *r, const uint32_t *a) { uint64_t c = 0; for (int i = 0; i < 8; i++) { c += (uint64_t)r[i] + a[i]; r[i] = (uint32_t)c; c >>= 32; } }
H2_HD void sub8(uint32_t *r, const uint32_t *a) { uint64_t b = 0; for (int i = 0; i < 8; i++) { uint64_t d = (uint64_t)r[i] - a[i] - b; r[i] = (uint32_t)d; b = (d >> 63) & 1; } }
// two's-complement 256-bit value -> magnitude (8 limbs, upper limbs zero for |v| < 2^160) and sign
H2_HD uint32_t abs8(uint32_t *r) {
    uint32_t neg = r[7] >> 31;
    if (neg) { uint64_t c = 1; for (int i = 0; i < 8; i++) { c += (uint64_t)(~r[i]); r[i] = (uint32_t)c; c >>= 32; } }
    return neg;
}
}  // namespace glv

// k (canonical, 8 limbs) -> |k1|, |k2| as 8-limb arrays (< 2^127) and their signs
template <class P> H2_HD void glv_decompose(const uint32_t (&k)[8], uint32_t (&k1)[8], uint32_t &neg1, uint32_t (&k2)[8], uint32_t &neg2) {
    typedef GlvConst<P> C;
    uint32_t g1[9], g2[9], a1[4], nb1[4], a2[4], b2[4];
    for (int i = 0; i < 9; i++) { g1[i] = C::g1(i); g2[i] = C::g2(i); }
    for (int i = 0; i < 4; i++) { a1[i] = C::a1(i); nb1[i] = C::nb1(i); a2[i] = C::a2(i); b2[i] = C::b2(i); }
    uint32_t w1[6], w2[6], c1[5], c2[5], t[8];
    glv::mul_part<8, 9, 11, 6>(k, g1, w1);       // limbs 11..16 of k g1; bit 383 = bit 31 of limb 11
    glv::mul_part<8, 9, 11, 6>(k, g2, w2);
    glv::round_shift(w1, c1);                    // c1 = (k g1 + 2^383) >> 384
    glv::round_shift(w2, c2);
    for (int i = 0; i < 8; i++) k1[i] = k[i];
    glv::mul_part<5, 4, 0, 8>(c1, a1, t); glv::sub8(k1, t);
    glv::mul_part<5, 4, 0, 8>(c2, a2, t); glv::sub8(k1, t);
    glv::mul_part<5, 4, 0, 8>(c1, nb1, k2);
    glv::mul_part<5, 4, 0, 8>(c2, b2, t); glv::sub8(k2, t);
    neg1 = glv::abs8(k1);
    neg2 = glv::abs8(k2);
}
template <class P> H2_HD fe glv_zeta() { fe r; for (int i = 0; i < 8; i++) r.v[i] = GlvConst<P>::zeta_mont(i); return r; }

}  // namespace h2
