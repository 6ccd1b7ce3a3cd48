// K13: curve point (de)compression -- C::to_bytes / C::from_bytes of pasta_curves as specified in
// /root/reference/book/src/background/curves.md:203-240, the encoding of every point in a proof transcript
// (halo2_proofs/src/transcript.rs) and in Params::{write, read} (halo2_proofs/src/poly/commitment.rs:168-205):
//     Enc(x, y) = x as 32 bytes little-endian with the LSB of y in the top bit of the last byte;  Enc(identity) = 0.
// Decompression needs y = sqrt(x^3 + 5): Tonelli-Shanks over the 2^32-order subgroup (both Pasta fields have 2-adicity
// 32, book/src/background/fields.md), ~250 squarings for a^((T-1)/2) plus at most 32 * 31 / 2 in the correction loop.
#pragma once
#include "curve.cuh"

namespace h2 {

struct SqrtConst {
    fe root;           // ROOT_OF_UNITY = 5^T, T = (m - 1) >> 32, Montgomery form
    uint32_t e[8];     // (T - 1) / 2
};

template <class P> H2_HD fe fe_pow_limbs(const fe &a, const uint32_t (&e)[8]) {
    fe acc = fe_one<P>();
    for (int i = 7; i >= 0; i--)
        for (int b = 31; b >= 0; b--) {
            acc = fe_sqr_call<P>(acc);
            if ((e[i] >> b) & 1u) acc = fe_mul_call<P>(acc, a);
        }
    return acc;
}
// host-side constants of the field (cheap: one exponentiation)
template <class P> inline SqrtConst make_sqrt_const() {
    SqrtConst K;
    uint32_t t[8];
    for (int i = 0; i < 7; i++) t[i] = mod_limb<P>(i + 1);     // T = (m - 1) >> 32: m's limb 0 is 1
    t[7] = 0;
    fe five = fe_zero();
    five.v[0] = 5;
    K.root = fe_pow_limbs<P>(fe_to_mont<P>(five), t);
    t[0] &= ~1u;                                               // T is odd
    for (int i = 0; i < 8; i++) K.e[i] = (t[i] >> 1) | (i < 7 ? t[i + 1] << 31 : 0u);
    return K;
}
// a square root of a (Montgomery form); false when a is not a square
template <class P> H2_HD bool fe_sqrt(const fe &a, const SqrtConst &K, fe &out) {
    if (fe_is_zero(a)) { out = a; return true; }
    const fe one = fe_one<P>();
    fe w = fe_pow_limbs<P>(a, K.e);
    fe x = fe_mul_call<P>(a, w);       // a^((T+1)/2)
    fe b = fe_mul_call<P>(x, w);       // a^T
    fe z = K.root;
    uint32_t v = 32;
    while (!fe_eq(b, one)) {
        uint32_t k = 0;
        fe b2 = b;
        while (!fe_eq(b2, one)) {
            b2 = fe_sqr_call<P>(b2);
            if (++k == v) return false;
        }
        fe zz = z;
        for (uint32_t i = 0; i + k + 1 < v; i++) zz = fe_sqr_call<P>(zz);
        x = fe_mul_call<P>(x, zz);
        z = fe_sqr_call<P>(zz);
        b = fe_mul_call<P>(b, z);
        v = k;
    }
    out = x;
    return true;
}

H2_HD void note_bad(uint32_t *first_bad, uint64_t i) {
#ifdef __CUDA_ARCH__
    atomicMin(first_bad, (uint32_t)i);
#else
    if ((uint32_t)i < *first_bad) *first_bad = (uint32_t)i;
#endif
}

template <class P> struct Codec {
    // affine (canonical or Montgomery) -> 32 bytes
    static H2_HD void compress_body(const affine *in, int in_mont, fe *out, uint64_t n, uint64_t i) {
        if (i >= n) return;
        affine p;
        p.x = fe_load(&in[i].x); p.y = fe_load(&in[i].y);
        if (in_mont) { p.x = fe_from_mont<P>(p.x); p.y = fe_from_mont<P>(p.y); }   // (0, 0) stays (0, 0)
        fe e = p.x;
        e.v[7] |= (p.y.v[0] & 1u) << 31;
        fe_store(out + i, e);
    }
    // 32 bytes -> affine; an invalid encoding (x >= m, x = 0 with the sign bit, x^3 + 5 not a square) writes the identity and
    // records its index in *first_bad (C::from_bytes returns None; Params::read fails with io::Error)
    static H2_HD void decompress_body(const fe *in, affine *out, int out_mont, const SqrtConst &K, uint32_t *first_bad, uint64_t n, uint64_t i) {
        if (i >= n) return;
        fe x = fe_load(in + i);
        const uint32_t ysign = x.v[7] >> 31;
        x.v[7] &= 0x7fffffffu;
        affine r;
        r.x = fe_zero(); r.y = fe_zero();
        bool ok = true, ident = false;
        if (fe_is_zero(x)) { ident = true; ok = ysign == 0; }
        else {
            bool lt = false;                                   // x < m, most significant limb first
            for (int j = 7; j >= 0; j--) {
                const uint32_t mj = mod_limb<P>(j);
                if (x.v[j] != mj) { lt = x.v[j] < mj; break; }
            }
            ok = lt;
        }
        if (ok && !ident) {
            const fe xm = fe_to_mont<P>(x);
            const fe one = fe_one<P>();
            fe five = fe_add<P>(fe_dbl<P>(fe_dbl<P>(one)), one);
            fe rhs = fe_add<P>(fe_mul_call<P>(fe_sqr_call<P>(xm), xm), five);
            fe y;
            ok = fe_sqrt<P>(rhs, K, y);
            if (ok) {
                fe yc = fe_from_mont<P>(y);
                if ((yc.v[0] & 1u) != ysign) { y = fe_neg<P>(y); yc = fe_from_mont<P>(y); }
                r.x = out_mont ? xm : x;
                r.y = out_mont ? y : yc;
            }
        }
        if (!ok) note_bad(first_bad, i);
        fe_store(&out[i].x, r.x); fe_store(&out[i].y, r.y);
    }
};

#if defined(__CUDACC__)
template <class P> __global__ void __launch_bounds__(128) compress_kernel(const affine *in, int in_mont, fe *out, uint64_t n) {
    Codec<P>::compress_body(in, in_mont, out, n, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
template <class P> __global__ void __launch_bounds__(128) decompress_kernel(const fe *in, affine *out, int out_mont, const SqrtConst K, uint32_t *first_bad,
                                                                           uint64_t n) {
    Codec<P>::decompress_body(in, out, out_mont, K, first_bad, n, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
#endif

}  // namespace h2
