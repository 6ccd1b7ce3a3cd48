// Device field library for the Pasta fields Fp / Fq (K0 in SURVEY.md section 2.1).
//
// Replaces the limb arithmetic of the un-vendored crate pasta_curves 0.5.1 that the
// reference reaches through ff::Field / PrimeField (halo2_proofs/src/arithmetic.rs:4-10,
// used at :243-246, :289-292 and throughout Bucket::add_assign :37-46).
//
// Representation: Montgomery form, R = 2^256, 8 x 32-bit limbs held in registers, always
// fully reduced to [0, m).  Both moduli have the shape
//     m = 2^254 + t,  t < 2^126,  limb0 = 1, limbs 4..6 = 0, limb7 = 0x40000000
// so  -m^-1 mod 2^32 = 0xffffffff (the Montgomery quotient digit is a negation) and the
// q*m product needs 3 real 32x32 multiplies + one "multiply" by 2^30.
//
// Multiplication is CIOS with the product rows split into even/odd 64-bit columns so that
// every mad.lo.cc / madc.hi.cc pair maps onto one IMAD.WIDE.U32(.X) in SASS and the carry
// chains never ripple (the sliding window's top limb is fresh every iteration).
//
// The same source compiles for the host (g++, no nvcc) with the PTX carry-flag instructions
// emulated; tests/kernel_emul uses that build to check kernel logic without a GPU.  The
// product path is the sm_100a build only.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define H2_HD __host__ __device__ __forceinline__
#define H2_D __device__ __forceinline__
#else
#define H2_HD inline
#define H2_D inline
// host-only build (tests/kernel_emul): minimal stand-ins for the CUDA vector types
struct alignas(16) uint4 { uint32_t x, y, z, w; };
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r = {x, y, z, w}; return r; }
struct alignas(8) uint2 { uint32_t x, y; };
inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r = {x, y}; return r; }
#endif

namespace h2 {

// ------------------------------------------------------------------ carry-flag primitives
namespace ptx {
#if defined(__CUDA_ARCH__)
H2_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
H2_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
H2_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
H2_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
H2_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
H2_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
// funnel shifts (ALU pipe); written as asm so that ptxas does not turn the pair q << 30, q >> 2 back into an IMAD.WIDE
H2_D uint32_t shl30(uint32_t a) { uint32_t r; asm volatile("shf.l.clamp.b32 %0, %1, %2, 30;" : "=r"(r) : "r"(0u), "r"(a)); return r; }
H2_D uint32_t shr2(uint32_t a) { uint32_t r; asm volatile("shf.r.clamp.b32 %0, %1, %2, 2;" : "=r"(r) : "r"(a), "r"(0u)); return r; }
H2_D uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
H2_D uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
H2_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
H2_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
H2_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
H2_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#else
// Host emulation of the PTX CC.CF flag (one per thread).
static thread_local uint32_t cf_ = 0;
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; cf_ = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b + cf_; cf_ = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + cf_; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t d = (uint64_t)a - b; cf_ = (uint32_t)(d >> 63); return (uint32_t)d; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t d = (uint64_t)a - b - cf_; cf_ = (uint32_t)(d >> 63); return (uint32_t)d; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - cf_; }
inline uint32_t shl30(uint32_t a) { return a << 30; }
inline uint32_t shr2(uint32_t a) { return a >> 2; }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)mul_lo(a, b) + c; cf_ = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)mul_lo(a, b) + c + cf_; cf_ = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)mul_hi(a, b) + c + cf_; cf_ = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return mul_hi(a, b) + c + cf_; }
#endif
}  // namespace ptx

// ------------------------------------------------------------------ field parameters
// Values checked against the reference goldens in tests/test_oracle_golden.py
// (moduli: halo2_proofs/tests/plonk_api.rs:591-592).
struct FpParams {   // Pallas base field / Vesta scalar field
    static constexpr int ID = 0;
    static constexpr uint32_t M1 = 0x992d30edu, M2 = 0x094cf91bu, M3 = 0x224698fcu;
    static H2_HD uint32_t one(int i) {   // R mod m
        constexpr uint32_t v[8] = {0xfffffffdu, 0x34786d38u, 0xe41914adu, 0x992c350bu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0x3fffffffu};
        return v[i];
    }
    static H2_HD uint32_t r2(int i) {    // R^2 mod m
        constexpr uint32_t v[8] = {0x0000000fu, 0x8c78ecb3u, 0x8b0de0e7u, 0xd7d30dbdu, 0xc3c95d18u, 0x7797a99bu, 0x7b9cb714u, 0x096d41afu};
        return v[i];
    }
    static H2_HD uint32_t r3(int i) {    // R^3 mod m (fe_inv_gcd: the inverse of a Montgomery residue, back in Montgomery form)
        constexpr uint32_t v[8] = {0x3a9e10f9u, 0xf185a599u, 0x6ac5b1d1u, 0xf6a68f3bu, 0x353fd42cu, 0xdf8d1014u, 0x2d2d9910u, 0x2ae30922u};
        return v[i];
    }
};
struct FqParams {   // Vesta base field / Pallas scalar field
    static constexpr int ID = 1;
    static constexpr uint32_t M1 = 0x8c46eb21u, M2 = 0x0994a8ddu, M3 = 0x224698fcu;
    static H2_HD uint32_t one(int i) {
        constexpr uint32_t v[8] = {0xfffffffdu, 0x5b2b3e9cu, 0xe3420567u, 0x992c350bu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0x3fffffffu};
        return v[i];
    }
    static H2_HD uint32_t r2(int i) {
        constexpr uint32_t v[8] = {0x0000000fu, 0xfc9678ffu, 0x891a16e3u, 0x67bb433du, 0x04ccf590u, 0x7fae2310u, 0x7ccfdaa9u, 0x096d41afu};
        return v[i];
    }
    static H2_HD uint32_t r3(int i) {
        constexpr uint32_t v[8] = {0x249dae4cu, 0x008b421cu, 0xdba41326u, 0xe13bda50u, 0x8e15cb63u, 0x88fececbu, 0x6e6792c8u, 0x07dd97a0u};
        return v[i];
    }
};
static constexpr uint32_t H2_M7 = 0x40000000u;   // top limb of both moduli

template <class P> H2_HD uint32_t mod_limb(int i) {
    return i == 0 ? 1u : i == 1 ? P::M1 : i == 2 ? P::M2 : i == 3 ? P::M3 : i == 7 ? H2_M7 : 0u;
}

// ------------------------------------------------------------------ element type
struct alignas(16) fe {
    uint32_t v[8];
};

H2_HD fe fe_zero() { fe r; for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
template <class P> H2_HD fe fe_one() { fe r; for (int i = 0; i < 8; i++) r.v[i] = P::one(i); return r; }
template <class P> H2_HD fe fe_r2() { fe r; for (int i = 0; i < 8; i++) r.v[i] = P::r2(i); return r; }

H2_HD bool fe_is_zero(const fe &a) {
    uint32_t t = a.v[0];
    for (int i = 1; i < 8; i++) t |= a.v[i];
    return t == 0;
}
H2_HD bool fe_eq(const fe &a, const fe &b) {
    uint32_t t = a.v[0] ^ b.v[0];
    for (int i = 1; i < 8; i++) t |= a.v[i] ^ b.v[i];
    return t == 0;
}

// r = a - m if a >= m else a   (a < 2^256)
template <class P> H2_HD void fe_cond_sub_mod(fe &a) {
    uint32_t s[8];
    s[0] = ptx::sub_cc(a.v[0], 1u);
    s[1] = ptx::subc_cc(a.v[1], P::M1);
    s[2] = ptx::subc_cc(a.v[2], P::M2);
    s[3] = ptx::subc_cc(a.v[3], P::M3);
    s[4] = ptx::subc_cc(a.v[4], 0u);
    s[5] = ptx::subc_cc(a.v[5], 0u);
    s[6] = ptx::subc_cc(a.v[6], 0u);
    s[7] = ptx::subc_cc(a.v[7], H2_M7);
    uint32_t borrow = ptx::subc(0u, 0u);   // 0xffffffff if a < m
    for (int i = 0; i < 8; i++) a.v[i] = borrow ? a.v[i] : s[i];
}

template <class P> H2_HD fe fe_add(const fe &a, const fe &b) {
    fe r;
    r.v[0] = ptx::add_cc(a.v[0], b.v[0]);
    for (int i = 1; i < 7; i++) r.v[i] = ptx::addc_cc(a.v[i], b.v[i]);
    r.v[7] = ptx::addc(a.v[7], b.v[7]);    // a + b < 2m < 2^256
    fe_cond_sub_mod<P>(r);
    return r;
}
template <class P> H2_HD fe fe_dbl(const fe &a) { return fe_add<P>(a, a); }

template <class P> H2_HD fe fe_sub(const fe &a, const fe &b) {
    fe r;
    r.v[0] = ptx::sub_cc(a.v[0], b.v[0]);
    for (int i = 1; i < 8; i++) r.v[i] = ptx::subc_cc(a.v[i], b.v[i]);
    uint32_t mask = ptx::subc(0u, 0u);     // all ones if borrow
    r.v[0] = ptx::add_cc(r.v[0], mask & 1u);
    r.v[1] = ptx::addc_cc(r.v[1], mask & P::M1);
    r.v[2] = ptx::addc_cc(r.v[2], mask & P::M2);
    r.v[3] = ptx::addc_cc(r.v[3], mask & P::M3);
    r.v[4] = ptx::addc_cc(r.v[4], 0u);
    r.v[5] = ptx::addc_cc(r.v[5], 0u);
    r.v[6] = ptx::addc_cc(r.v[6], 0u);
    r.v[7] = ptx::addc(r.v[7], mask & H2_M7);
    return r;
}
template <class P> H2_HD fe fe_neg(const fe &a) {
    return fe_sub<P>(fe_zero(), a);
}

// ------------------------------------------------------------------ Montgomery multiplication
// CIOS on an even/odd split accumulator:
//   ev[k] sits at limb position k      (pairs (0,1)(2,3)(4,5)(6,7))
//   od[k] sits at limb position k + 1  (pairs (1,2)(3,4)(5,6)(7,8))
// One iteration adds a * bi, then q * m with q = -ev[0] (leaving ev[0] == 0), then the window
// slides down one limb.  The iteration is cut into four carry chains (product odd/even,
// reduction odd/even) so that fe_mul2 can interleave the chains of two INDEPENDENT
// multiplications in program order: ptxas overlaps adjacent independent chains but does not
// reorder across whole multiplications, and a lone warp (the MSM's serial tails) otherwise
// runs at ~0.36 IPC.
namespace mont {
using namespace ptx;
// a * bi, odd columns.  CARRY_IN: the caller has just issued the add.cc that folds the
// left-over limb into ev[0]; its carry enters at position 1 = od[0].
template <bool FIRST> H2_HD void prod_od(uint32_t (&od)[8], const fe &a, uint32_t bi) {
    if (FIRST) {
        for (int k = 0; k < 4; k++) { od[2 * k] = mul_lo(a.v[2 * k + 1], bi); od[2 * k + 1] = mul_hi(a.v[2 * k + 1], bi); }
    } else {
        od[0] = madc_lo_cc(a.v[1], bi, od[0]);
        od[1] = madc_hi_cc(a.v[1], bi, od[1]);
        od[2] = madc_lo_cc(a.v[3], bi, od[2]);
        od[3] = madc_hi_cc(a.v[3], bi, od[3]);
        od[4] = madc_lo_cc(a.v[5], bi, od[4]);
        od[5] = madc_hi_cc(a.v[5], bi, od[5]);
        od[6] = madc_lo_cc(a.v[7], bi, od[6]);
        od[7] = madc_hi(a.v[7], bi, od[7]);          // no carry out: od <= A / 2^32 < 2^256
    }
}
template <bool FIRST> H2_HD void prod_ev(uint32_t (&ev)[8], uint32_t (&od)[8], const fe &a, uint32_t bi) {
    if (FIRST) {
        for (int k = 0; k < 4; k++) { ev[2 * k] = mul_lo(a.v[2 * k], bi); ev[2 * k + 1] = mul_hi(a.v[2 * k], bi); }
    } else {
        ev[0] = mad_lo_cc(a.v[0], bi, ev[0]);
        ev[1] = madc_hi_cc(a.v[0], bi, ev[1]);
        ev[2] = madc_lo_cc(a.v[2], bi, ev[2]);
        ev[3] = madc_hi_cc(a.v[2], bi, ev[3]);
        ev[4] = madc_lo_cc(a.v[4], bi, ev[4]);
        ev[5] = madc_hi_cc(a.v[4], bi, ev[5]);
        ev[6] = madc_lo_cc(a.v[6], bi, ev[6]);
        ev[7] = madc_hi_cc(a.v[6], bi, ev[7]);
        od[7] = addc(od[7], 0u);                     // carry out of position 7 lands on position 8
    }
}
// q * m, odd columns: m1 @1, m3 @3, 2^30 @7.  CIN: a carry into position 1 is pending in CC (fe_sqr's
// product-free iterations, where no prod_od chain has consumed the fold-in carry).
template <class P, bool CIN = false> H2_HD void red_od(uint32_t (&od)[8], uint32_t q) {
    od[0] = CIN ? madc_lo_cc(q, P::M1, od[0]) : mad_lo_cc(q, P::M1, od[0]);
    od[1] = madc_hi_cc(q, P::M1, od[1]);
    od[2] = madc_lo_cc(q, P::M3, od[2]);
    od[3] = madc_hi_cc(q, P::M3, od[3]);
    od[4] = addc_cc(od[4], 0u);
    od[5] = addc_cc(od[5], 0u);
    od[6] = addc_cc(od[6], shl30(q));                // q * 2^30 with shifts: keeps 2 of 8 multiplies per round
    od[7] = addc(od[7], shr2(q));                    // off the IMAD pipe, the one that bounds the MSM
}
// q * m, even columns: 1 @0, m2 @2
template <class P> H2_HD void red_ev(uint32_t (&ev)[8], uint32_t (&od)[8], uint32_t q) {
    ev[0] = add_cc(ev[0], q);                        // == 0, carry = (old ev[0] != 0)
    ev[1] = addc_cc(ev[1], 0u);
    ev[2] = madc_lo_cc(q, P::M2, ev[2]);
    ev[3] = madc_hi_cc(q, P::M2, ev[3]);
    ev[4] = addc_cc(ev[4], 0u);
    ev[5] = addc_cc(ev[5], 0u);
    ev[6] = addc_cc(ev[6], 0u);
    ev[7] = addc_cc(ev[7], 0u);
    od[7] = addc(od[7], 0u);
}
// window slide: position 0 (== 0) drops out, ev <- od, od <- ev >> 2 limbs; returns the
// left-over limb (new position 0) that the next iteration folds into its ev[0].
H2_HD uint32_t slide(uint32_t (&x)[8]) {
    uint32_t left = x[1];
    for (int k = 0; k < 6; k++) x[k] = x[k + 2];
    x[6] = 0; x[7] = 0;
    return left;
}
// result = ev (positions 0..7) + od (positions 1..8, od[7] == 0 since the value is < 2m);
// the carry of the last fold-in add is still pending in CC.
template <class P> H2_HD fe finish(const uint32_t (&x)[8], const uint32_t (&y)[8]) {
    fe r;
    r.v[0] = x[0];
    r.v[1] = addc_cc(x[1], y[0]);
    r.v[2] = addc_cc(x[2], y[1]);
    r.v[3] = addc_cc(x[3], y[2]);
    r.v[4] = addc_cc(x[4], y[3]);
    r.v[5] = addc_cc(x[5], y[4]);
    r.v[6] = addc_cc(x[6], y[5]);
    r.v[7] = addc(x[7], y[6]);
    return r;
}
// one iteration; (ev, od) on entry, roles swapped on exit
template <class P, bool FIRST> H2_HD void iter(uint32_t (&ev)[8], uint32_t (&od)[8], const fe &a, uint32_t bi, uint32_t left) {
    if (!FIRST) ev[0] = add_cc(ev[0], left);
    prod_od<FIRST>(od, a, bi);
    prod_ev<FIRST>(ev, od, a, bi);
    uint32_t q = 0u - ev[0];
    red_od<P>(od, q);
    red_ev<P>(ev, od, q);
}
}  // namespace mont

template <class P> H2_HD fe fe_mul(const fe &a, const fe &b) {
    uint32_t x[8], y[8], left;
    mont::iter<P, true>(x, y, a, b.v[0], 0u);  left = mont::slide(x);
    mont::iter<P, false>(y, x, a, b.v[1], left); left = mont::slide(y);
    mont::iter<P, false>(x, y, a, b.v[2], left); left = mont::slide(x);
    mont::iter<P, false>(y, x, a, b.v[3], left); left = mont::slide(y);
    mont::iter<P, false>(x, y, a, b.v[4], left); left = mont::slide(x);
    mont::iter<P, false>(y, x, a, b.v[5], left); left = mont::slide(y);
    mont::iter<P, false>(x, y, a, b.v[6], left); left = mont::slide(x);
    mont::iter<P, false>(y, x, a, b.v[7], left); left = mont::slide(y);
    x[0] = ptx::add_cc(x[0], left);
    fe r = mont::finish<P>(x, y);
    fe_cond_sub_mod<P>(r);
    return r;
}

// Two independent products r0 = a0 * b0, r1 = a1 * b1 with their carry chains interleaved.
template <class P, bool FIRST>
H2_HD void mont_iter2(uint32_t (&ev0)[8], uint32_t (&od0)[8], const fe &a0, uint32_t b0, uint32_t l0,
                      uint32_t (&ev1)[8], uint32_t (&od1)[8], const fe &a1, uint32_t b1, uint32_t l1) {
    using namespace mont;
    if (!FIRST) ev0[0] = ptx::add_cc(ev0[0], l0);
    prod_od<FIRST>(od0, a0, b0);
    if (!FIRST) ev1[0] = ptx::add_cc(ev1[0], l1);
    prod_od<FIRST>(od1, a1, b1);
    prod_ev<FIRST>(ev0, od0, a0, b0);
    prod_ev<FIRST>(ev1, od1, a1, b1);
    uint32_t q0 = 0u - ev0[0], q1 = 0u - ev1[0];
    red_od<P>(od0, q0);
    red_od<P>(od1, q1);
    red_ev<P>(ev0, od0, q0);
    red_ev<P>(ev1, od1, q1);
}
template <class P> H2_HD void fe_mul2(fe &r0, const fe &a0, const fe &b0, fe &r1, const fe &a1, const fe &b1) {
    uint32_t x0[8], y0[8], x1[8], y1[8], l0, l1;
    mont_iter2<P, true>(x0, y0, a0, b0.v[0], 0u, x1, y1, a1, b1.v[0], 0u);
    l0 = mont::slide(x0); l1 = mont::slide(x1);
#define H2_IT2(EV0, OD0, EV1, OD1, K)                                                  \
    mont_iter2<P, false>(EV0, OD0, a0, b0.v[K], l0, EV1, OD1, a1, b1.v[K], l1);        \
    l0 = mont::slide(EV0); l1 = mont::slide(EV1);
    H2_IT2(y0, x0, y1, x1, 1)
    H2_IT2(x0, y0, x1, y1, 2)
    H2_IT2(y0, x0, y1, x1, 3)
    H2_IT2(x0, y0, x1, y1, 4)
    H2_IT2(y0, x0, y1, x1, 5)
    H2_IT2(x0, y0, x1, y1, 6)
    H2_IT2(y0, x0, y1, x1, 7)
#undef H2_IT2
    x0[0] = ptx::add_cc(x0[0], l0);
    fe t0 = mont::finish<P>(x0, y0);
    x1[0] = ptx::add_cc(x1[0], l1);
    fe t1 = mont::finish<P>(x1, y1);
    fe_cond_sub_mod<P>(t0);
    fe_cond_sub_mod<P>(t1);
    r0 = t0; r1 = t1;
}
// Dedicated squaring: the 28 off-diagonal products once (rows of IMAD.WIDE carry chains into an even- and an
// odd-position accumulator, E and O, indexed by absolute limb position), doubled, plus the 8 diagonal squares:
// 36 wide multiplies instead of 64.  The low half then runs through the same sliding-window reduction as fe_mul
// with no products to add, and the high half is added at the end:
//   (T + Q m) / 2^256 = T_hi + (T_lo + Q m) / 2^256 <= T_hi + m < 2m      (T_hi < m^2 / 2^256 < m / 2).
template <class P> H2_HD fe fe_sqr(const fe &a) {
    using namespace ptx;
    const uint32_t a0 = a.v[0], a1 = a.v[1], a2 = a.v[2], a3 = a.v[3], a4 = a.v[4], a5 = a.v[5], a6 = a.v[6], a7 = a.v[7];
    uint32_t E[14], O[15];
    // row 0 opens both accumulators
    O[1] = mul_lo(a0, a1); O[2] = mul_hi(a0, a1); O[3] = mul_lo(a0, a3); O[4] = mul_hi(a0, a3);
    O[5] = mul_lo(a0, a5); O[6] = mul_hi(a0, a5); O[7] = mul_lo(a0, a7); O[8] = mul_hi(a0, a7);
    E[2] = mul_lo(a0, a2); E[3] = mul_hi(a0, a2); E[4] = mul_lo(a0, a4); E[5] = mul_hi(a0, a4);
    E[6] = mul_lo(a0, a6); E[7] = mul_hi(a0, a6);
    // row 1
    O[3] = mad_lo_cc(a1, a2, O[3]); O[4] = madc_hi_cc(a1, a2, O[4]); O[5] = madc_lo_cc(a1, a4, O[5]); O[6] = madc_hi_cc(a1, a4, O[6]);
    O[7] = madc_lo_cc(a1, a6, O[7]); O[8] = madc_hi_cc(a1, a6, O[8]); O[9] = addc(0u, 0u);
    E[4] = mad_lo_cc(a1, a3, E[4]); E[5] = madc_hi_cc(a1, a3, E[5]); E[6] = madc_lo_cc(a1, a5, E[6]); E[7] = madc_hi_cc(a1, a5, E[7]);
    E[8] = madc_lo_cc(a1, a7, 0u); E[9] = madc_hi(a1, a7, 0u);
    // row 2
    O[5] = mad_lo_cc(a2, a3, O[5]); O[6] = madc_hi_cc(a2, a3, O[6]); O[7] = madc_lo_cc(a2, a5, O[7]); O[8] = madc_hi_cc(a2, a5, O[8]);
    O[9] = madc_lo_cc(a2, a7, O[9]); O[10] = madc_hi(a2, a7, 0u);
    E[6] = mad_lo_cc(a2, a4, E[6]); E[7] = madc_hi_cc(a2, a4, E[7]); E[8] = madc_lo_cc(a2, a6, E[8]); E[9] = madc_hi_cc(a2, a6, E[9]);
    E[10] = addc(0u, 0u);
    // row 3
    O[7] = mad_lo_cc(a3, a4, O[7]); O[8] = madc_hi_cc(a3, a4, O[8]); O[9] = madc_lo_cc(a3, a6, O[9]); O[10] = madc_hi_cc(a3, a6, O[10]);
    O[11] = addc(0u, 0u);
    E[8] = mad_lo_cc(a3, a5, E[8]); E[9] = madc_hi_cc(a3, a5, E[9]); E[10] = madc_lo_cc(a3, a7, E[10]); E[11] = madc_hi(a3, a7, 0u);
    // row 4
    O[9] = mad_lo_cc(a4, a5, O[9]); O[10] = madc_hi_cc(a4, a5, O[10]); O[11] = madc_lo_cc(a4, a7, O[11]); O[12] = madc_hi(a4, a7, 0u);
    E[10] = mad_lo_cc(a4, a6, E[10]); E[11] = madc_hi_cc(a4, a6, E[11]); E[12] = addc(0u, 0u);
    // row 5
    O[11] = mad_lo_cc(a5, a6, O[11]); O[12] = madc_hi_cc(a5, a6, O[12]); O[13] = addc(0u, 0u);
    E[12] = mad_lo_cc(a5, a7, E[12]); E[13] = madc_hi(a5, a7, 0u);
    // row 6
    O[13] = mad_lo_cc(a6, a7, O[13]); O[14] = madc_hi(a6, a7, 0u);
    // U = E + O (positions 1..14; U < 2^479 so nothing leaves limb 14)
    uint32_t U[15];
    U[1] = O[1];
    U[2] = add_cc(E[2], O[2]);
    for (int p = 3; p <= 13; p++) U[p] = addc_cc(E[p], O[p]);
    U[14] = addc(O[14], 0u);
    // T = 2 U + sum_i a_i^2 2^(64 i)
    uint32_t T[16];
    T[0] = mul_lo(a0, a0);
    T[1] = add_cc((U[1] << 1), mul_hi(a0, a0));
    for (int i = 1; i < 8; i++) {
        const uint32_t ai = a.v[i];
        T[2 * i] = addc_cc((U[2 * i] << 1) | (U[2 * i - 1] >> 31), mul_lo(ai, ai));
        if (i < 7) T[2 * i + 1] = addc_cc((U[2 * i + 1] << 1) | (U[2 * i] >> 31), mul_hi(ai, ai));
        else T[15] = addc(U[14] >> 31, mul_hi(ai, ai));
    }
    // Montgomery reduction of T[0..8) with the sliding window of fe_mul (ev = x, od = y, roles swap each round)
    uint32_t x[8], y[8], left, q;
    for (int k = 0; k < 8; k++) { x[k] = T[k]; y[k] = 0u; }
    q = 0u - x[0]; mont::red_od<P>(y, q); mont::red_ev<P>(x, y, q); left = mont::slide(x);
#define H2_RED(EV, OD)                                                                   \
    EV[0] = add_cc(EV[0], left); q = 0u - EV[0];                                         \
    mont::red_od<P, true>(OD, q); mont::red_ev<P>(EV, OD, q); left = mont::slide(EV);
    H2_RED(y, x) H2_RED(x, y) H2_RED(y, x) H2_RED(x, y) H2_RED(y, x) H2_RED(x, y) H2_RED(y, x)
#undef H2_RED
    x[0] = add_cc(x[0], left);
    fe r = mont::finish<P>(x, y);
    r.v[0] = add_cc(r.v[0], T[8]);
    for (int k = 1; k < 7; k++) r.v[k] = addc_cc(r.v[k], T[8 + k]);
    r.v[7] = addc(r.v[7], T[15]);
    fe_cond_sub_mod<P>(r);
    return r;
}

// canonical <-> Montgomery
template <class P> H2_HD fe fe_to_mont(const fe &a) { return fe_mul<P>(a, fe_r2<P>()); }
template <class P> H2_HD fe fe_from_mont(const fe &a) {
    fe one = fe_zero(); one.v[0] = 1u;
    return fe_mul<P>(a, one);
}

// a^(m-2) by square-and-multiply over the fixed exponent (m - 2 = 2^254 + t - 2).
template <class P> H2_HD fe fe_inv(const fe &a) {
    fe acc = fe_one<P>();
    // exponent limbs of m - 2 (limb0 = 0xffffffff after the borrow from limb1)
    const uint32_t e[8] = {0xffffffffu, P::M1 - 1u, P::M2, P::M3, 0u, 0u, 0u, H2_M7};
    for (int i = 7; i >= 0; i--) {
        for (int b = 31; b >= 0; b--) {
            acc = fe_sqr<P>(acc);
            if ((e[i] >> b) & 1u) acc = fe_mul<P>(acc, a);
        }
    }
    return acc;
}

// ------------------------------------------------------------------ inversion by divsteps (Bernstein-Yang "safegcd")
// The Fermat ladder above is ~380 DEPENDENT multiplies: 0.1 ms for a lone thread (every batch_normalize of a handful of
// commitments, every IPA round that returns affine points) and ~300 multiply-equivalents of issue slots when every lane
// inverts (the shared inversion of a batch of affine additions).  This one runs 20 rounds of 30 division steps on the low
// words of (f, g) -- 32-bit scalar work off the multiply pipe -- each followed by one 2x2 matrix update of the full-width
// pairs (f, g) and (d, e), in 9 signed 30-bit limbs: ~11 k instructions, ~1.7 k of them wide multiplies (a Montgomery
// multiply is 244 / 63), no data-dependent branch (uniform across a warp), ~10x shorter as a dependent chain.
// 600 division steps cover any 256-bit odd modulus (the published bound for this variant is 590); g = 0 yields 0 like
// fe_inv.  Moduli here are = 1 mod 2^30, so the "modulus^-1 mod 2^30" of the (d, e) update is 1.
struct s30 { int32_t v[9]; };                    // sum v[i] 2^(30 i), limbs signed
H2_HD s30 s30_from_fe(const fe &a) {
    s30 r;
    for (int i = 0; i < 9; i++) {
        const int bit = 30 * i, j = bit >> 5, sh = bit & 31;
        uint32_t w = a.v[j] >> sh;
        if (sh > 2 && j + 1 < 8) w |= a.v[j + 1] << (32 - sh);
        r.v[i] = (int32_t)(w & 0x3fffffffu);
    }
    return r;
}
H2_HD fe s30_to_fe(const s30 &a) {               // limbs in [0, 2^30), value < 2^256
    fe r;
    for (int j = 0; j < 8; j++) {
        const int bit = 32 * j, i = bit / 30, sh = bit % 30;      // sh <= 14: two limbs cover a word
        r.v[j] = ((uint32_t)a.v[i] >> sh) | ((uint32_t)a.v[i + 1] << (30 - sh));
    }
    return r;
}
template <class P> H2_HD s30 s30_modulus() {
    fe m;
    for (int i = 0; i < 8; i++) m.v[i] = mod_limb<P>(i);
    return s30_from_fe(m);
}
// 30 division steps on the low words; returns the new zeta = -(delta + 1/2) and the transition matrix (u v; q r), entries
// in [-2^30, 2^30], with  2^30 (f', g') = (u v; q r) (f, g)
H2_HD int32_t divsteps30(int32_t zeta, uint32_t f, uint32_t g, int32_t (&t)[4]) {
    uint32_t u = 1, v = 0, q = 0, r = 1;
    for (int i = 0; i < 30; i++) {
        uint32_t m1 = (uint32_t)(zeta >> 31), m2 = 0u - (g & 1u);
        const uint32_t x = (f ^ m1) - m1, y = (u ^ m1) - m1, z = (v ^ m1) - m1;     // -(f, u, v) when zeta < 0
        g += x & m2; q += y & m2; r += z & m2;
        m1 &= m2;                                                                    // zeta < 0 and g odd: swap roles
        zeta = (zeta ^ (int32_t)m1) - 1;
        f += g & m1; u += q & m1; v += r & m1;
        g >>= 1; u <<= 1; v <<= 1;
    }
    t[0] = (int32_t)u; t[1] = (int32_t)v; t[2] = (int32_t)q; t[3] = (int32_t)r;
    return zeta;
}
// (f, g) <- (u v; q r) (f, g) / 2^30   (exact)
H2_HD void s30_update_fg(s30 &f, s30 &g, const int32_t (&t)[4]) {
    const int64_t u = t[0], v = t[1], q = t[2], r = t[3];
    int64_t cf = u * f.v[0] + v * g.v[0], cg = q * f.v[0] + r * g.v[0];
    cf >>= 30; cg >>= 30;
    for (int i = 1; i < 9; i++) {
        cf += u * f.v[i] + v * g.v[i];
        cg += q * f.v[i] + r * g.v[i];
        f.v[i - 1] = (int32_t)cf & 0x3fffffff; cf >>= 30;
        g.v[i - 1] = (int32_t)cg & 0x3fffffff; cg >>= 30;
    }
    f.v[8] = (int32_t)cf; g.v[8] = (int32_t)cg;
}
// (d, e) <- (u v; q r) (d, e) / 2^30 mod m: a multiple of m is added first so that the division is exact
template <class P> H2_HD void s30_update_de(s30 &d, s30 &e, const int32_t (&t)[4], const s30 &m) {
    const int32_t u = t[0], v = t[1], q = t[2], r = t[3];
    const int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;
    int32_t md = (u & sd) + (v & se), me = (q & sd) + (r & se);
    int64_t cd = (int64_t)u * d.v[0] + (int64_t)v * e.v[0], ce = (int64_t)q * d.v[0] + (int64_t)r * e.v[0];
    md -= (int32_t)(((uint32_t)cd + (uint32_t)md) & 0x3fffffffu);      // m^-1 mod 2^30 = 1
    me -= (int32_t)(((uint32_t)ce + (uint32_t)me) & 0x3fffffffu);
    cd += (int64_t)m.v[0] * md; ce += (int64_t)m.v[0] * me;
    cd >>= 30; ce >>= 30;
    for (int i = 1; i < 9; i++) {
        cd += (int64_t)u * d.v[i] + (int64_t)v * e.v[i];
        ce += (int64_t)q * d.v[i] + (int64_t)r * e.v[i];
        if (i < 5 || i == 8) { cd += (int64_t)m.v[i] * md; ce += (int64_t)m.v[i] * me; }   // limbs 5..7 of both moduli are 0
        d.v[i - 1] = (int32_t)cd & 0x3fffffff; cd >>= 30;
        e.v[i - 1] = (int32_t)ce & 0x3fffffff; ce >>= 30;
    }
    d.v[8] = (int32_t)cd; e.v[8] = (int32_t)ce;
}
// a^-1 for a Montgomery residue a (0 -> 0), result in Montgomery form
template <class P> H2_HD fe fe_inv_gcd(const fe &a) {
    const s30 m = s30_modulus<P>();
    s30 d, e, f = m, g = s30_from_fe(a);
    for (int i = 0; i < 9; i++) { d.v[i] = 0; e.v[i] = 0; }
    e.v[0] = 1;
    int32_t zeta = -1;
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (int it = 0; it < 20; it++) {
        int32_t t[4];
        zeta = divsteps30(zeta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
        s30_update_de<P>(d, e, t, m);
        s30_update_fg(f, g, t);
    }
    // f = +-1 and d = +- a^-1 in (-2m, m): add m if negative, negate if f < 0, add m again if still negative
    const int32_t M30 = 0x3fffffff;
    int32_t add = d.v[8] >> 31, neg = f.v[8] >> 31;
    for (int i = 0; i < 9; i++) d.v[i] = ((d.v[i] + (m.v[i] & add)) ^ neg) - neg;
    for (int i = 0; i < 8; i++) { d.v[i + 1] += d.v[i] >> 30; d.v[i] &= M30; }
    add = d.v[8] >> 31;
    for (int i = 0; i < 9; i++) d.v[i] += m.v[i] & add;
    for (int i = 0; i < 8; i++) { d.v[i + 1] += d.v[i] >> 30; d.v[i] &= M30; }
    fe r3;
    for (int i = 0; i < 8; i++) r3.v[i] = P::r3(i);
    return fe_mul<P>(s30_to_fe(d), r3);          // (a R)^-1 R^3 / R = a^-1 R
}

// 128-bit global/shared memory access helpers
H2_HD fe fe_load(const fe *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 lo = q[0], hi = q[1];
    fe r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
}
H2_HD void fe_store(fe *p, const fe &a) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}

}  // namespace h2
