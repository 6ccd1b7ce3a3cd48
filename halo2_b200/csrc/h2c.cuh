// K16: hash_to_curve -- C::CurveExt::hash_to_curve(domain_prefix)(message), the generator derivation of Params::new
// (/root/reference/halo2_proofs/src/poly/commitment.rs:46-58 for g[i], :102-105 for w and u).  The reference calls the
// un-vendored crate pasta_curves 0.5.1; what that computes is the hash-to-curve suite "<curve>_XMD:BLAKE2b_SSWU_RO_"
// of RFC 9380:
//   1. hash_to_field: expand_message_xmd (section 5.3.1) over BLAKE2b-512 (128-byte blocks, no key, all-zero
//      personalisation) with DST = domain_prefix || "-" || curve || "_XMD:BLAKE2b_SSWU_RO_", 128 output bytes, each half
//      read big-endian and reduced modulo the coordinate field -> u0, u1;
//   2. simplified SWU (section 6.6.2) of u0 and u1 onto iso-<curve>: y^2 = x^3 + A x + 1265, Z = -13, sgn0 = parity;
//   3. the two images added on the iso curve; 4. the 3-isogeny iso-<curve> -> <curve> (section 6.6.3).
// Constants: the iso curve is the codomain of Velu's 3-isogeny from y^2 = x^3 + 5 whose kernel has x0^3 = -20
// (A = -30 x0^2, B = 1265); the map back is the dual isogeny, kernel x = xk:
//     x' = (x + T/(x - xk) + U/(x - xk)^2) / 9,   y' = y (1 - T/(x - xk)^2 - 2U/(x - xk)^3) / 27,
//     T = 6 xk^2 + 2A,  U = 4 (xk^3 + A xk + B).
// Only A and xk are literals (derived from Velu's formulas by the test suite's big-integer checker, and pinned -- like this
// file, through tests/ -- on the reference's golden commitments, tests/plonk_api.rs:958-982); everything else is computed from
// them on the host, with the check that the dual lands on y^2 = x^3 + 5 (A - 5T = 0, B - 7(U + xk T) = 729 * 5).
// One thread per message: 4 BLAKE2b compressions for short messages and ~3 500 field multiplies (two inv0, three square
// roots on average, two affine-formula inversions) -- a one-off set-up cost next to the EC-FFT that follows it.
#pragma once
#include "codec.cuh"

namespace h2 {
#define cm fe_mul_call
#define cs fe_sqr_call

// ------------------------------------------------------------------------------------------------ BLAKE2b-512
struct Blake2b {
    uint64_t h[8];
    uint64_t t;            // bytes compressed so far (messages here are < 2^64 bytes)
    uint8_t buf[128];
    uint32_t fill;
};
H2_HD uint64_t b2_iv(int i) {
    constexpr uint64_t IV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
    return IV[i];
}
H2_HD uint32_t b2_sigma(int r, int i) {
    constexpr uint8_t S[10][16] = {{0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
                                   {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
                                   {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
                                   {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
                                   {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
    return S[r][i];
}
H2_HD uint64_t b2_rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
H2_HD void b2_init(Blake2b &S) {
    for (int i = 0; i < 8; i++) S.h[i] = b2_iv(i);
    S.h[0] ^= 0x01010000ull ^ 64ull;   // digest length 64, no key, fanout = depth = 1; salt and personalisation zero
    S.t = 0;
    S.fill = 0;
}
H2_HD void b2_compress(Blake2b &S, bool last) {
    uint64_t m[16], v[16];
    for (int i = 0; i < 16; i++) {
        uint64_t w = 0;
        for (int j = 7; j >= 0; j--) w = (w << 8) | S.buf[8 * i + j];
        m[i] = w;
    }
    for (int i = 0; i < 8; i++) { v[i] = S.h[i]; v[8 + i] = b2_iv(i); }
    v[12] ^= S.t;
    if (last) v[14] = ~v[14];
    for (int r = 0; r < 12; r++) {
        const int rr = r % 10;
#define H2_B2G(a, b, c, d, i)                                   \
    v[a] = v[a] + v[b] + m[b2_sigma(rr, 2 * (i))];              \
    v[d] = b2_rotr(v[d] ^ v[a], 32);                            \
    v[c] = v[c] + v[d];                                         \
    v[b] = b2_rotr(v[b] ^ v[c], 24);                            \
    v[a] = v[a] + v[b] + m[b2_sigma(rr, 2 * (i) + 1)];          \
    v[d] = b2_rotr(v[d] ^ v[a], 16);                            \
    v[c] = v[c] + v[d];                                         \
    v[b] = b2_rotr(v[b] ^ v[c], 63);
        H2_B2G(0, 4, 8, 12, 0) H2_B2G(1, 5, 9, 13, 1) H2_B2G(2, 6, 10, 14, 2) H2_B2G(3, 7, 11, 15, 3)
        H2_B2G(0, 5, 10, 15, 4) H2_B2G(1, 6, 11, 12, 5) H2_B2G(2, 7, 8, 13, 6) H2_B2G(3, 4, 9, 14, 7)
#undef H2_B2G
    }
    for (int i = 0; i < 8; i++) S.h[i] ^= v[i] ^ v[8 + i];
}
H2_HD void b2_update(Blake2b &S, const uint8_t *p, uint32_t len) {
    for (uint32_t i = 0; i < len; i++) {
        if (S.fill == 128) {            // a full buffer is only compressed once more input follows it
            S.t += 128;
            b2_compress(S, false);
            S.fill = 0;
        }
        S.buf[S.fill++] = p[i];
    }
}
H2_HD void b2_update_byte(Blake2b &S, uint8_t b) { b2_update(S, &b, 1); }
H2_HD void b2_final(Blake2b &S, uint8_t (&out)[64]) {
    S.t += S.fill;
    for (uint32_t i = S.fill; i < 128; i++) S.buf[i] = 0;
    b2_compress(S, true);
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(S.h[i] >> (8 * j));
}

// ------------------------------------------------------------------------------------------------ constants
template <class P> struct H2cLit;   // keyed by the curve's COORDINATE field (FpParams = Pallas); canonical limbs
template <> struct H2cLit<FpParams> {
    static constexpr uint32_t A[8] = {0x657a014bu, 0x92bb4b0bu, 0x1a27a59fu, 0xb7413458u, 0x58370742u, 0x49be2d72u, 0xb0ea8c9cu, 0x18354a2eu};
    static constexpr uint32_t XK[8] = {0x4ba19471u, 0x6a57031bu, 0x1ff0c7cdu, 0x4301a71du, 0x8fdb5ac3u, 0x52cfc019u, 0x11fb3180u, 0x115468c1u};
    static constexpr const char *NAME = "pallas";
};
template <> struct H2cLit<FqParams> {
    static constexpr uint32_t A[8] = {0x42eaa6b1u, 0xc515ad72u, 0x7d01b212u, 0x9673928cu, 0x96f78773u, 0x81639c4du, 0xe592271au, 0x267f9b2eu};
    static constexpr uint32_t XK[8] = {0x286f2e8cu, 0xea8f4dd1u, 0x6fef5204u, 0xbf4c98bdu, 0xd251d4a6u, 0x75d5c33au, 0x54bf6d15u, 0x1ae90dbdu};
    static constexpr const char *NAME = "vesta";
};

struct H2cConst {           // Montgomery form
    fe A, B, Z, mBA, BZA;   // iso curve, SWU Z = -13, -B/A, B/(Z A)
    fe xk, T, U, inv9, inv27;
    fe r3;                  // R^3 mod m: to_mont of the high half of a 512-bit value
    SqrtConst sq;
    uint8_t dst[256];       // DST || len(DST)
    uint32_t dst_len;       // including the length byte
    bool ok;
};

template <class P> H2_HD fe fe_small(uint32_t x) {
    fe r = fe_zero();
    r.v[0] = x;
    return fe_to_mont<P>(r);
}
// host: all constants of one (curve, domain prefix)
template <class P> inline H2cConst make_h2c_const(const char *domain_prefix) {
    H2cConst K;
    fe a, xk;
    for (int i = 0; i < 8; i++) { a.v[i] = H2cLit<P>::A[i]; xk.v[i] = H2cLit<P>::XK[i]; }
    K.A = fe_to_mont<P>(a);
    K.xk = fe_to_mont<P>(xk);
    K.B = fe_small<P>(1265);
    K.Z = fe_neg<P>(fe_small<P>(13));
    const fe ainv = fe_inv<P>(K.A);
    K.mBA = fe_neg<P>(fe_mul<P>(K.B, ainv));
    K.BZA = fe_mul<P>(K.B, fe_inv<P>(fe_mul<P>(K.Z, K.A)));
    const fe xk2 = fe_sqr<P>(K.xk);
    K.T = fe_add<P>(fe_mul<P>(fe_small<P>(6), xk2), fe_dbl<P>(K.A));
    K.U = fe_mul<P>(fe_small<P>(4), fe_add<P>(fe_add<P>(fe_mul<P>(xk2, K.xk), fe_mul<P>(K.A, K.xk)), K.B));
    K.inv9 = fe_inv<P>(fe_small<P>(9));
    K.inv27 = fe_inv<P>(fe_small<P>(27));
    K.r3 = fe_mul<P>(fe_r2<P>(), fe_r2<P>());
    K.sq = make_sqrt_const<P>();
    // the dual isogeny must land on y^2 = x^3 + 5: A - 5T = 0 and B - 7 (U + xk T) = 3^6 * 5
    const fe W = fe_add<P>(K.U, fe_mul<P>(K.xk, K.T));
    K.ok = fe_eq(K.A, fe_mul<P>(fe_small<P>(5), K.T)) &&
           fe_eq(fe_sub<P>(K.B, fe_mul<P>(fe_small<P>(7), W)), fe_small<P>(729 * 5));
    uint32_t n = 0;
    auto put = [&](const char *s) { for (; *s && n < 255; s++) K.dst[n++] = (uint8_t)*s; };
    put(domain_prefix); put("-"); put(H2cLit<P>::NAME); put("_XMD:BLAKE2b_SSWU_RO_");
    K.dst[n] = (uint8_t)n;
    K.dst_len = n + 1;
    if (n >= 255) K.ok = false;
    return K;
}

// ------------------------------------------------------------------------------------------------ the map
// 64 big-endian bytes -> field element (Montgomery): F::from_uniform_bytes of the reversed digest
template <class P> H2_HD fe fe_from_be64(const uint8_t (&d)[64], const H2cConst &K) {
    fe lo, hi;
    for (int i = 0; i < 8; i++) {
        uint32_t l = 0, h = 0;
        for (int j = 3; j >= 0; j--) {                 // limb i of lo = bytes 63-4i-3 .. 63-4i (most significant first)
            l = (l << 8) | d[63 - 4 * i - j];
            h = (h << 8) | d[31 - 4 * i - j];
        }
        lo.v[i] = l; hi.v[i] = h;
    }
    for (int k = 0; k < 3; k++) { fe_cond_sub_mod<P>(lo); fe_cond_sub_mod<P>(hi); }   // < 2^256 < 4m
    return fe_add<P>(fe_mul_call<P>(lo, fe_r2<P>()), fe_mul_call<P>(hi, K.r3));      // lo R + hi R^2 = (lo + hi 2^256) R
}
template <class P> H2_HD fe fe_inv_call(const fe &a) {          // a^(m-2); 0 -> 0 (inv0)
    uint32_t e[8];
    for (int i = 0; i < 8; i++) e[i] = mod_limb<P>(i);
    e[0] = 0xffffffffu; e[1] -= 1u;                             // m - 2: limb 0 is 1, limb 1 is non-zero for both moduli
    return fe_pow_limbs<P>(a, e);
}
template <class P> H2_HD uint32_t fe_parity(const fe &a_mont) { return fe_from_mont<P>(a_mont).v[0] & 1u; }

// simplified SWU onto the iso curve (affine, never the identity)
template <class P> H2_HD void h2c_swu(const fe &u, const H2cConst &K, fe &x, fe &y) {
    const fe one = fe_one<P>();
    const fe zu2 = cm<P>(K.Z, cs<P>(u));
    const fe ta = fe_add<P>(cs<P>(zu2), zu2);
    const fe tv1 = fe_inv_call<P>(ta);
    fe x1 = fe_is_zero(tv1) ? K.BZA : cm<P>(K.mBA, fe_add<P>(one, tv1));
    const fe gx1 = fe_add<P>(cm<P>(fe_add<P>(cs<P>(x1), K.A), x1), K.B);
    if (fe_sqrt<P>(gx1, K.sq, y)) x = x1;
    else {
        x = cm<P>(zu2, x1);
        const fe gx2 = fe_add<P>(cm<P>(fe_add<P>(cs<P>(x), K.A), x), K.B);
        fe_sqrt<P>(gx2, K.sq, y);                       // exactly one of gx1, gx2 is a square (Z is a non-residue)
    }
    if (fe_parity<P>(u) != fe_parity<P>(y)) y = fe_neg<P>(y);
}
// hash one message to an affine point of the curve (Montgomery; identity = (0, 0))
template <class P> H2_HD affine h2c_point(const uint8_t *msg, uint32_t msg_len, const H2cConst &K) {
    uint8_t b0[64], b1[64], b2[64];
    Blake2b S;
    b2_init(S);
    for (int i = 0; i < 128; i++) S.buf[i] = 0;         // Z_pad: one zero block
    S.fill = 128;
    b2_update(S, msg, msg_len);
    b2_update_byte(S, 0); b2_update_byte(S, 128); b2_update_byte(S, 0);   // I2OSP(128, 2) || I2OSP(0, 1)
    b2_update(S, K.dst, K.dst_len);
    b2_final(S, b0);
    b2_init(S);
    b2_update(S, b0, 64); b2_update_byte(S, 1); b2_update(S, K.dst, K.dst_len);
    b2_final(S, b1);
    b2_init(S);
    for (int i = 0; i < 64; i++) b2_update_byte(S, b0[i] ^ b1[i]);
    b2_update_byte(S, 2); b2_update(S, K.dst, K.dst_len);
    b2_final(S, b2);

    fe x1, y1, x2, y2;
    h2c_swu<P>(fe_from_be64<P>(b1, K), K, x1, y1);
    h2c_swu<P>(fe_from_be64<P>(b2, K), K, x2, y2);

    affine out; out.x = fe_zero(); out.y = fe_zero();
    // q0 + q1 on the iso curve
    fe num, den;
    if (fe_eq(x1, x2)) {
        if (!fe_eq(y1, y2) || fe_is_zero(y1)) return out;            // q1 = -q0
        num = fe_add<P>(fe_add<P>(fe_dbl<P>(cs<P>(x1)), cs<P>(x1)), K.A);
        den = fe_dbl<P>(y1);
    } else { num = fe_sub<P>(y2, y1); den = fe_sub<P>(x2, x1); }
    const fe lam = cm<P>(num, fe_inv_call<P>(den));
    const fe x3 = fe_sub<P>(fe_sub<P>(cs<P>(lam), x1), x2);
    const fe y3 = fe_sub<P>(cm<P>(lam, fe_sub<P>(x1, x3)), y1);
    // the 3-isogeny
    const fe d = fe_sub<P>(x3, K.xk);
    if (fe_is_zero(d)) return out;                                   // a kernel point
    const fe di = fe_inv_call<P>(d);
    const fe di2 = cs<P>(di);
    const fe udi2 = cm<P>(K.U, di2);
    out.x = cm<P>(fe_add<P>(fe_add<P>(x3, cm<P>(K.T, di)), udi2), K.inv9);
    const fe f = fe_sub<P>(fe_sub<P>(fe_one<P>(), cm<P>(K.T, di2)), fe_dbl<P>(cm<P>(udi2, di)));
    out.y = cm<P>(cm<P>(y3, f), K.inv27);
    return out;
}
// messages: n x msg_len bytes, or -- gen_params != 0 -- the generator messages of Params::new: 0 || i as u32 LE
// (poly/commitment.rs:54-56) for i = first + index
template <class P> H2_HD void h2c_body(const uint8_t *msgs, uint32_t msg_len, int gen_params, uint64_t first, const H2cConst &K, affine *out,
                                       int out_mont, uint64_t n, uint64_t i) {
    if (i >= n) return;
    uint8_t gm[5];
    const uint8_t *m = msgs + i * msg_len;
    if (gen_params) {
        const uint32_t idx = (uint32_t)(first + i);
        gm[0] = 0; gm[1] = (uint8_t)idx; gm[2] = (uint8_t)(idx >> 8); gm[3] = (uint8_t)(idx >> 16); gm[4] = (uint8_t)(idx >> 24);
        m = gm; msg_len = 5;
    }
    affine p = h2c_point<P>(m, msg_len, K);
    if (!out_mont) { p.x = fe_from_mont<P>(p.x); p.y = fe_from_mont<P>(p.y); }
    fe_store(&out[i].x, p.x); fe_store(&out[i].y, p.y);
}

#if defined(__CUDACC__)
template <class P> __global__ void __launch_bounds__(64) h2c_kernel(const uint8_t *msgs, uint32_t msg_len, int gen_params, uint64_t first,
                                                                   const H2cConst K, affine *out, int out_mont, uint64_t n) {
    h2c_body<P>(msgs, msg_len, gen_params, first, K, out, out_mont, n, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
#endif

#undef cm
#undef cs
}  // namespace h2
