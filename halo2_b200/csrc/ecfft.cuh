// K10 / K11: best_fft with G = curve point, and batch_normalize.
//
// K10 restates the butterfly network of best_fft (/root/reference/halo2_proofs/src/arithmetic.rs:192-295) at
// G = C::Curve -- the FftGroup bound (:17-27) makes group_scale a SCALAR MULTIPLICATION and group_add / group_sub point
// additions.  The reference has one call site, Params::new (poly/commitment.rs:77-94): g_lagrange is the inverse EC-FFT
// of g, every output then multiplied by 2^-k (:84-89) and batch-normalised (:91-101).  Same network as ntt.cuh (bit
// reversal :207-212, twiddles w^i :215-221, log n radix-2 DIT stages), no step assumes w^n = 1.
//
// Layout: the points live as XYZZ (128 B, Montgomery) at their network POSITION p = bitrev(j) from the load on; stage s
// (1-based) pairs p and p + 2^(s-1) inside blocks of 2^s with twiddle w^((p mod 2^(s-1)) * 2^(log_n - s)) -- one work
// item per butterfly, log n launches.  The butterfly's cost is the scalar multiplication t = tw * b: the twiddle is split
// by the GLV endomorphism (glv.cuh, |k1|, |k2| < 2^127) and t = k1 b + k2 phi(b) is one joint double-and-add over the table
// {±b, ±phi(b), ±b ± phi(b)}: 127 doublings + <= 127 additions instead of 255 + 255.  Butterflies with twiddle exponent 0
// (the whole first stage) skip the multiplication (w^0 = 1: same group element as the reference's `* 1`).
//
// K11 batch_normalize (group::Curve::batch_normalize; plonk/prover.rs:99,311, poly/commitment.rs:65,95): Montgomery's
// trick in chunks of H2_NORM_CHUNK points per thread -- one field inversion per chunk, identity points skipped and
// written as (0, 0).
#pragma once
#include "curve.cuh"
#include "glv.cuh"
#include "ntt.cuh"
#include "msm.cuh"   // ld_affine / st_affine / st_jacobian

namespace h2 {

#define H2_NORM_CHUNK 16

H2_HD jacobian ld_jac(const jacobian *p) {
    jacobian r;
    r.x = fe_load(&p->x); r.y = fe_load(&p->y); r.z = fe_load(&p->z);
    return r;
}
template <class P> H2_HD xyzz xyzz_from_jacobian(const jacobian &j) {
    if (fe_is_zero(j.z)) return xyzz_identity();
    xyzz t;
    t.x = j.x; t.y = j.y;
    t.zz = fe_sqr_call<P>(j.z);
    t.zzz = fe_mul_call<P>(t.zz, j.z);
    return t;
}

// k * b for a projective b and a canonical scalar k < r (8 limbs): GLV halves, joint double-and-add, most significant
// bit first.  One call site for the addition (the operand is SELECTED, then added), so divergent lanes of a warp do not
// serialise three copies of the group law.
template <class P> H2_HD xyzz xyzz_scalar_mul_glv(const xyzz &b, const uint32_t (&k)[8]) {
    if (xyzz_is_identity(b)) return b;
    uint32_t k1[8], k2[8], n1, n2;
    glv_decompose<P>(k, k1, n1, k2, n2);
    xyzz tab[3];                       // [0] = ±b, [1] = ±phi(b), [2] = their sum
    tab[0] = b;
    if (n1) xyzz_neg<P>(tab[0]);
    tab[1] = b;
    tab[1].x = fe_mul_call<P>(b.x, glv_zeta<P>());
    if (n2) xyzz_neg<P>(tab[1]);
    tab[2] = tab[0];
    xyzz_add<P>(tab[2], tab[1]);
    xyzz acc = xyzz_identity();
    for (int bit = 126; bit >= 0; bit--) {
        xyzz_double<P>(acc);
        uint32_t sel = ((k1[bit >> 5] >> (bit & 31)) & 1u) | (((k2[bit >> 5] >> (bit & 31)) & 1u) << 1);
        if (sel) {
            const xyzz &op = tab[sel - 1];
            xyzz_add<P>(acc, op);
        }
    }
    return acc;
}

// ---- quad-cooperative group law -------------------------------------------------------------------------------------
// A stage of the EC-FFT at k = 14 has 8192 butterflies, each a serial chain of ~2500 field multiplications: with one
// thread per butterfly the GPU is latency-bound (a lone warp needs ~0.36 us per multiply).  Four consecutive lanes (a
// QUAD) therefore share one butterfly: every level of the formulas below is four INDEPENDENT products, one per lane,
// exchanged by shuffle (quad_mul4).  XYZZ doubling is 9 products in 3 levels, the general addition 14 in 4, and the
// butterfly's (a + t, a - t) pair -- which shares u1, u2, s1, pp and therefore ZZ3, ZZZ3 -- 16 in 4 instead of 28 in 8.
//
// Control flow is WARP-UNIFORM by construction: identity operands and skipped additions are handled by selects, the
// rare P = ±Q case by a per-lane serial fallback behind a warp vote (every lane of a quad holds full copies of its
// operands, so any lane can finish alone).  All 32 lanes therefore reach every shuffle together and the shuffles use the
// constant full mask -- with a per-quad runtime mask ptxas wraps each of the 32 shuffles of a level in a
// MATCH / WARPSYNC / BSSY sequence, and with divergent quads the bit loop was 72 KB of code per iteration: a lone warp
// then runs at the instruction-fetch rate (first version, measured on B200: 19k cycles per scalar bit).  The multiply is
// ONE out-of-line body (fe_mul_call) for the same reason.
// On the host (tests/kernel_emul) quad_mul4 is four plain multiplies and the votes are the lane's own flag, so the
// formulas and the select logic are checked against the oracle there; only the shuffle itself is device-only.
H2_HD bool warp_any(bool p) {
#ifdef __CUDA_ARCH__
    return __any_sync(0xffffffffu, p) != 0;
#else
    return p;
#endif
}
H2_HD fe fe_pick(bool c, const fe &a, const fe &b) {
    fe r;
    for (int i = 0; i < 8; i++) r.v[i] = c ? a.v[i] : b.v[i];
    return r;
}
H2_HD xyzz xyzz_pick(bool c, const xyzz &a, const xyzz &b) {
    xyzz r;
    r.x = fe_pick(c, a.x, b.x); r.y = fe_pick(c, a.y, b.y); r.zz = fe_pick(c, a.zz, b.zz); r.zzz = fe_pick(c, a.zzz, b.zzz);
    return r;
}
template <class P>
H2_HD void quad_mul4(const fe &a0, const fe &b0, const fe &a1, const fe &b1, const fe &a2, const fe &b2, const fe &a3, const fe &b3,
                     fe &r0, fe &r1, fe &r2, fe &r3) {
#ifdef __CUDA_ARCH__
    const uint32_t sub = threadIdx.x & 3u;
    fe x, y;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        x.v[i] = sub == 0 ? a0.v[i] : sub == 1 ? a1.v[i] : sub == 2 ? a2.v[i] : a3.v[i];
        y.v[i] = sub == 0 ? b0.v[i] : sub == 1 ? b1.v[i] : sub == 2 ? b2.v[i] : b3.v[i];
    }
    fe r = fe_mul_call<P>(x, y);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r0.v[i] = __shfl_sync(0xffffffffu, r.v[i], 0, 4);
        r1.v[i] = __shfl_sync(0xffffffffu, r.v[i], 1, 4);
        r2.v[i] = __shfl_sync(0xffffffffu, r.v[i], 2, 4);
        r3.v[i] = __shfl_sync(0xffffffffu, r.v[i], 3, 4);
    }
#else
    r0 = fe_mul<P>(a0, b0); r1 = fe_mul<P>(a1, b1); r2 = fe_mul<P>(a2, b2); r3 = fe_mul<P>(a3, b3);
#endif
}
// acc = 2 acc (dbl-2008-s-1, a = 0) in 3 levels.  Branch-free: the identity (0, 0, 0, 0) maps to itself.
template <class P> H2_HD void xyzz_double_q(xyzz &a) {
    fe u = fe_dbl<P>(a.y), v, xx, w, s, mm, zz3, t0, t1, zzz3, d0, d1;
    quad_mul4<P>(u, u, a.x, a.x, u, u, a.x, a.x, v, xx, d0, d1);
    fe m = fe_add<P>(fe_dbl<P>(xx), xx);
    quad_mul4<P>(u, v, a.x, v, m, m, v, a.zz, w, s, mm, zz3);
    fe x3 = fe_sub<P>(fe_sub<P>(mm, s), s);
    quad_mul4<P>(m, fe_sub<P>(s, x3), w, a.y, w, a.zzz, w, a.zzz, t0, t1, zzz3, d0);
    a.x = x3; a.y = fe_sub<P>(t0, t1); a.zz = zz3; a.zzz = zzz3;
}
// sum = a + b and (when DIFF) diff = a - b (add-2008-s) in 4 levels; a - b shares everything but r = s2 - s1.
// Must be reached by all lanes of the warp together (see above).
template <class P, bool DIFF> H2_HD void xyzz_addsub_q(const xyzz &a, const xyzz &b, xyzz &sum, xyzz &diff) {
    const bool a_id = xyzz_is_identity(a), b_id = xyzz_is_identity(b);
    fe u1, u2, s1, s2;
    quad_mul4<P>(a.x, b.zz, b.x, a.zz, a.y, b.zzz, b.y, a.zzz, u1, u2, s1, s2);
    fe pp = fe_sub<P>(u2, u1);
    fe r = fe_sub<P>(s2, s1);
    fe rn = fe_neg<P>(fe_add<P>(s2, s1));                  // -s2 - s1: the r of a + (-b)
    const bool degenerate = !a_id && !b_id && fe_is_zero(pp);   // b = ±a
    fe pp2, rr, rrn, zz12, ppp, q, zz3, zzz12, t1, t1n, t2, zzz3;
    quad_mul4<P>(pp, pp, r, r, rn, rn, a.zz, b.zz, pp2, rr, rrn, zz12);
    quad_mul4<P>(pp, pp2, u1, pp2, zz12, pp2, a.zzz, b.zzz, ppp, q, zz3, zzz12);
    fe x3 = fe_sub<P>(fe_sub<P>(fe_sub<P>(rr, ppp), q), q);
    fe x3n = fe_sub<P>(fe_sub<P>(fe_sub<P>(rrn, ppp), q), q);
    quad_mul4<P>(r, fe_sub<P>(q, x3), rn, fe_sub<P>(q, x3n), s1, ppp, zzz12, ppp, t1, t1n, t2, zzz3);
    xyzz g;
    g.x = x3; g.y = fe_sub<P>(t1, t2); g.zz = zz3; g.zzz = zzz3;
    sum = xyzz_pick(b_id, a, xyzz_pick(a_id, b, g));
    if (DIFF) {
        xyzz nb = b;
        xyzz_neg<P>(nb);
        g.x = x3n; g.y = fe_sub<P>(t1n, t2);
        diff = xyzz_pick(b_id, a, xyzz_pick(a_id, nb, g));
    }
    if (warp_any(degenerate)) {                            // rare; serial per lane, no shuffles inside
        if (degenerate) {
            xyzz dbl = a;
            xyzz_double<P>(dbl);
            const bool same = fe_is_zero(r);
            sum = same ? dbl : xyzz_identity();
            if (DIFF) diff = same ? xyzz_identity() : dbl;
        }
    }
}
// xyzz_scalar_mul_glv for a quad: 3 + 4 levels per bit instead of 9 + 14 multiplications.  Warp-uniform: the addition of
// a bit runs when ANY quad of the warp has a non-zero digit pair (with 8 quads that is 99.99 % of the bits) and quads with
// a zero pair discard its result.
template <class P> H2_HD xyzz xyzz_scalar_mul_glv_q(const xyzz &b, const uint32_t (&k)[8]) {
    uint32_t k1[8], k2[8], n1, n2;
    glv_decompose<P>(k, k1, n1, k2, n2);
    xyzz tab[3], unused;
    tab[0] = b;
    if (n1) xyzz_neg<P>(tab[0]);
    tab[1] = b;
    tab[1].x = fe_mul_call<P>(b.x, glv_zeta<P>());
    if (n2) xyzz_neg<P>(tab[1]);
    xyzz_addsub_q<P, false>(tab[0], tab[1], tab[2], unused);
    xyzz acc = xyzz_identity();
    for (int bit = 126; bit >= 0; bit--) {
        xyzz_double_q<P>(acc);
        const uint32_t sel = ((k1[bit >> 5] >> (bit & 31)) & 1u) | (((k2[bit >> 5] >> (bit & 31)) & 1u) << 1);
        if (warp_any(sel != 0)) {
            xyzz op = xyzz_pick(sel == 2, tab[1], xyzz_pick(sel == 3, tab[2], tab[0])), res;
            xyzz_addsub_q<P, false>(acc, op, res, unused);
            acc = xyzz_pick(sel != 0, res, acc);
        }
    }
    return acc;
}
// the quad's lanes each store one coordinate (device); the host emulation stores all four
H2_HD void st_xyzz_q(xyzz *p, const xyzz &a) {
#ifdef __CUDA_ARCH__
    const uint32_t sub = threadIdx.x & 3u;
    fe_store(sub == 0 ? &p->x : sub == 1 ? &p->y : sub == 2 ? &p->zz : &p->zzz, fe_pick(sub == 0, a.x, fe_pick(sub == 1, a.y, fe_pick(sub == 2, a.zz, a.zzz))));
#else
    st_xyzz(p, a);
#endif
}

// P = coordinate field of the curve, PS = its scalar field (the field of the twiddles)
template <class P, class PS> struct EcFft {
    // input j -> work[bitrev(j)]; Jacobian (pasta's Ep/Eq layout) or affine input
    static H2_HD void load_jac_body(const jacobian *in, int canonical, xyzz *work, uint32_t log_n, uint64_t j) {
        jacobian p = ld_jac(in + j);
        if (canonical) { p.x = fe_to_mont<P>(p.x); p.y = fe_to_mont<P>(p.y); p.z = fe_to_mont<P>(p.z); }
        st_xyzz(work + bitrev32((uint32_t)j, log_n), xyzz_from_jacobian<P>(p));
    }
    static H2_HD void load_affine_body(const affine *in, int canonical, xyzz *work, uint32_t log_n, uint64_t j) {
        affine p = ld_affine(in + j);
        if (canonical && !affine_is_identity(p)) { p.x = fe_to_mont<P>(p.x); p.y = fe_to_mont<P>(p.y); }
        st_xyzz(work + bitrev32((uint32_t)j, log_n), xyzz_from_affine<P>(p));
    }
    // butterfly t of stage s (1-based): arithmetic.rs:237-249 / :276-293
    static H2_HD void stage_body(xyzz *work, const fe *tw, uint32_t log_n, uint32_t s, uint64_t t) {
        const uint64_t half = 1ull << (s - 1);
        const uint64_t i = t & (half - 1);
        const uint64_t ia = ((t >> (s - 1)) << s) + i, ib = ia + half;
        xyzz a = ld_xyzz(work + ia), b = ld_xyzz(work + ib);
        if (i) {
            fe w = fe_from_mont<PS>(fe_load(tw + (i << (log_n - s))));
            b = xyzz_scalar_mul_glv<P>(b, w.v);
        }
        xyzz d = a, nb = b;
        xyzz_neg<P>(nb);
        xyzz_add<P>(a, b);
        xyzz_add<P>(d, nb);
        st_xyzz(work + ia, a);
        st_xyzz(work + ib, d);
    }
    // the same butterfly run by a quad of lanes (t = quad index).  Called by EVERY lane of the warp: quads past the end
    // (`active` false) compute on butterfly 0 and store nothing; the scalar multiplication runs for the whole warp as
    // soon as one of its quads has a twiddle exponent != 0 (w^0 = 1 goes through the ladder unchanged).
    static H2_HD void stage_body_q(xyzz *work, const fe *tw, uint32_t log_n, uint32_t s, uint64_t t, bool active) {
        if (!active) t = 0;
        const uint64_t half = 1ull << (s - 1);
        const uint64_t i = t & (half - 1);
        const uint64_t ia = ((t >> (s - 1)) << s) + i, ib = ia + half;
        xyzz a = ld_xyzz(work + ia), b = ld_xyzz(work + ib);
        if (warp_any(i != 0)) {
            fe w = fe_from_mont<PS>(fe_load(tw + (i << (log_n - s))));
            b = xyzz_scalar_mul_glv_q<P>(b, w.v);
        }
        xyzz sum, diff;
        xyzz_addsub_q<P, true>(a, b, sum, diff);
        if (active) {
            st_xyzz_q(work + ia, sum);
            st_xyzz_q(work + ib, diff);
        }
    }
    static H2_HD void scale_body_q(xyzz *work, const fe &scale_canon, uint64_t i, bool active) {
        xyzz p = ld_xyzz(work + (active ? i : 0));
        p = xyzz_scalar_mul_glv_q<P>(p, scale_canon.v);
        if (active) st_xyzz_q(work + i, p);
    }
    // `*g *= scale` (poly/commitment.rs:84-89); scale canonical
    static H2_HD void scale_body(xyzz *work, const fe &scale_canon, uint64_t i) {
        xyzz p = ld_xyzz(work + i);
        st_xyzz(work + i, xyzz_scalar_mul_glv<P>(p, scale_canon.v));
    }
    static H2_HD void store_jac_body(const xyzz *work, jacobian *out, int canonical, uint64_t i) {
        jacobian r = xyzz_to_jacobian<P>(ld_xyzz(work + i));
        if (canonical) { r.x = fe_from_mont<P>(r.x); r.y = fe_from_mont<P>(r.y); r.z = fe_from_mont<P>(r.z); }
        st_jacobian(out + i, r);
    }
};

// K11.  Thread t normalises points [t * H2_NORM_CHUNK, ...).  Input either XYZZ (Montgomery) or Jacobian.
template <class P> struct Normalize {
    static H2_HD jacobian get(const xyzz *in_xyzz, const jacobian *in_jac, int in_canonical, uint64_t i) {
        if (in_xyzz) return xyzz_to_jacobian<P>(ld_xyzz(in_xyzz + i));
        jacobian p = ld_jac(in_jac + i);
        if (in_canonical) { p.x = fe_to_mont<P>(p.x); p.y = fe_to_mont<P>(p.y); p.z = fe_to_mont<P>(p.z); }
        return p;
    }
    static H2_HD void body(const xyzz *in_xyzz, const jacobian *in_jac, int in_canonical, affine *out, int out_canonical, uint64_t n,
                           uint64_t t) {
        const uint64_t lo = t * H2_NORM_CHUNK;
        if (lo >= n) return;
        const uint32_t m = (uint32_t)(n - lo < H2_NORM_CHUNK ? n - lo : H2_NORM_CHUNK);
        fe pre[H2_NORM_CHUNK];
        fe acc = fe_one<P>();
        for (uint32_t i = 0; i < m; i++) {
            jacobian p = get(in_xyzz, in_jac, in_canonical, lo + i);
            pre[i] = acc;
            if (!fe_is_zero(p.z)) acc = fe_mul_call<P>(acc, p.z);
        }
        acc = fe_inv_gcd<P>(acc);
        for (uint32_t i = m; i-- > 0;) {
            jacobian p = get(in_xyzz, in_jac, in_canonical, lo + i);
            affine r;
            if (fe_is_zero(p.z)) { r.x = fe_zero(); r.y = fe_zero(); }
            else {
                fe zi = fe_mul_call<P>(acc, pre[i]);
                acc = fe_mul_call<P>(acc, p.z);
                fe zi2 = fe_sqr_call<P>(zi);
                r.x = fe_mul_call<P>(p.x, zi2);
                r.y = fe_mul_call<P>(p.y, fe_mul_call<P>(zi2, zi));
                if (out_canonical) { r.x = fe_from_mont<P>(r.x); r.y = fe_from_mont<P>(r.y); }
            }
            st_affine(out + lo + i, r);
        }
    }
};

#if defined(__CUDACC__)
template <class P, class PS> __global__ void __launch_bounds__(128) ecfft_load_jac_kernel(const jacobian *in, int canonical, xyzz *work, uint32_t log_n) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < (1ull << log_n)) EcFft<P, PS>::load_jac_body(in, canonical, work, log_n, j);
}
template <class P, class PS> __global__ void __launch_bounds__(128) ecfft_load_affine_kernel(const affine *in, int canonical, xyzz *work, uint32_t log_n) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < (1ull << log_n)) EcFft<P, PS>::load_affine_body(in, canonical, work, log_n, j);
}
// 64-thread CTAs: at k = 14 a stage has only 8192 butterflies, so small CTAs spread them over all SMs
template <class P, class PS> __global__ void __launch_bounds__(64) ecfft_stage_kernel(xyzz *work, const fe *tw, uint32_t log_n, uint32_t s) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (1ull << (log_n - 1))) EcFft<P, PS>::stage_body(work, tw, log_n, s, t);
}
template <class P, class PS> __global__ void __launch_bounds__(64) ecfft_scale_kernel(xyzz *work, fe scale_canon, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) EcFft<P, PS>::scale_body(work, scale_canon, i);
}
// quad forms: 4 lanes per butterfly / point, 64-thread CTAs = 16 quads.  Every lane runs the body (uniform control flow).
template <class P, class PS> __global__ void __launch_bounds__(64) ecfft_stage_quad_kernel(xyzz *work, const fe *tw, uint32_t log_n, uint32_t s) {
    uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    EcFft<P, PS>::stage_body_q(work, tw, log_n, s, t, t < (1ull << (log_n - 1)));
}
template <class P, class PS> __global__ void __launch_bounds__(64) ecfft_scale_quad_kernel(xyzz *work, fe scale_canon, uint64_t n) {
    uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    EcFft<P, PS>::scale_body_q(work, scale_canon, i, i < n);
}
template <class P, class PS> __global__ void __launch_bounds__(128) ecfft_store_jac_kernel(const xyzz *work, jacobian *out, int canonical, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) EcFft<P, PS>::store_jac_body(work, out, canonical, i);
}
template <class P> __global__ void __launch_bounds__(64) normalize_kernel(const xyzz *in_xyzz, const jacobian *in_jac, int in_canonical, affine *out,
                                                                         int out_canonical, uint64_t n) {
    Normalize<P>::body(in_xyzz, in_jac, in_canonical, out, out_canonical, n, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
#endif

}  // namespace h2
