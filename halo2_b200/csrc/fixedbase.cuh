// K12: direct-sum fixed-base MSM for SMALL resident generator sets (k <= 15: the prover's commits and IPA rounds).
//
// Same result as best_multiexp (/root/reference/halo2_proofs/src/arithmetic.rs:143-180) on (scalars, g): the group
// element sum_i k_i G_i.  What changes is the decomposition.  The bucket method (msm.cuh) spends a k = 14 commit almost
// entirely in LATENCY: sort, work items, accumulate, three bucket-reduce kernels, window combine -- ~20 dependent
// launches of 10-70 us for 40 us worth of multiplies.  The generators of a Params never change
// (poly/commitment.rs:26-33), so for small sets the table can hold every digit multiple instead of every window shift:
//
//     D[i][w][m] = m * 2^(8 w) * G_i      w < 32 windows of 8 bits, m = 1 .. 128 (signed digits), affine
//
// (a generator's 4096 entries are contiguous: the 32 gathers of one scalar stay inside one 256 KiB region and the lanes of a
// warp inside a few MB -- with the generator index innermost every gather of a warp hit a different 2 MB page of the
// 4.3 GB table and both the build and the accumulation ran at the page-walk rate)
//
// 32 x 128 x 64 B = 256 KiB per generator -- 4.3 GB at k = 14, nothing next to 180 GB of HBM -- and a commit is a
// plain SUM of the n x 32 table entries the signed base-256 digits select: no buckets, no sort, no reduce chain.
//   fb_accum_kernel    one thread per (scalar, slice of 32 / split windows): recode, gather (64 B, random), mixed add
//   fb_reduce_kernel   64 quads per CTA: each quad adds f partial sums, then a shared-memory tree; repeated until one
//                      point per set is left (2 launches at k = 14), the last one converts to the Jacobian result
// Three launches instead of ~20.  `sets` independent scalar vectors (a batch of polynomials, or the L / R pair of an IPA
// round) go through the same launches.
#pragma once
#include "ecfft.cuh"
#include "msm.cuh"

namespace h2 {

#define H2_FB_BITS 8
#define H2_FB_WINDOWS 32
#define H2_FB_MULTIPLES 128
#define H2_FB_NORM 16          // multiples per batch inversion in the table build
#define H2_FB_QUADS 64         // quads per CTA of the reduce
#define H2_FB_MAX_FAN 8        // partial sums a quad adds before the tree

struct FbPlan {
    uint64_t total;        // scalars per set (<= the set's registered points)
    uint32_t sets;         // scalars are laid out [set][total], results [set]
    uint32_t split;        // threads per scalar: each takes 32 / split windows (1, 2, 4, 8)
    uint32_t scalars_mont;
};

template <class P, class PS> struct FixedBase {
    // Table build, one thread per (window w, generator i): the 128 multiples of B = wtab[w][i] = 2^(8 w) G_i by repeated
    // mixed addition, brought back to affine 16 at a time (XYZZ coordinates parked in the destination slots, one
    // inversion per 16 points).
    static H2_HD void table_body(const affine *wtab, affine *dtab, uint64_t count, uint64_t stride, uint64_t t) {
        if (t >= (uint64_t)H2_FB_WINDOWS * count) return;
        const uint64_t w = t / count, i = t % count;
        const affine B = ld_affine(wtab + w * stride + i);
        affine *dst = dtab + (i * H2_FB_WINDOWS + w) * H2_FB_MULTIPLES;   // multiple m lives at dst[m - 1]
        if (affine_is_identity(B)) {
            for (uint32_t m = 0; m < H2_FB_MULTIPLES; m++) st_affine(dst + m, B);
            return;
        }
        xyzz R = xyzz_from_affine<P>(B);
        for (uint32_t c0 = 0; c0 < H2_FB_MULTIPLES; c0 += H2_FB_NORM) {
            fe zz[H2_FB_NORM], zzz[H2_FB_NORM], pre[H2_FB_NORM];
            fe run = fe_one<P>();
            for (uint32_t j = 0; j < H2_FB_NORM; j++) {
                const uint32_t m = c0 + j + 1;
                if (m == 2) R = xyzz_double_affine<P>(B);
                else if (m > 2) xyzz_add_mixed<P>(R, B);               // m * B != identity, != B: the group order is prime
                affine park; park.x = R.x; park.y = R.y;
                st_affine(dst + (m - 1), park);
                zz[j] = R.zz; zzz[j] = R.zzz; pre[j] = run;
                run = fe_mul_call<P>(run, fe_mul_call<P>(R.zz, R.zzz));
            }
            fe inv = fe_inv_gcd<P>(run);
            for (uint32_t j = H2_FB_NORM; j-- > 0;) {
                fe id = fe_mul_call<P>(inv, pre[j]);                    // 1 / (zz zzz)
                inv = fe_mul_call<P>(inv, fe_mul_call<P>(zz[j], zzz[j]));
                affine a = ld_affine(dst + (c0 + j));
                a.x = fe_mul_call<P>(a.x, fe_mul_call<P>(id, zzz[j]));   // X / ZZ
                a.y = fe_mul_call<P>(a.y, fe_mul_call<P>(id, zz[j]));    // Y / ZZZ
                st_affine(dst + (c0 + j), a);
            }
        }
    }

    // Thread u = ((set * split) + part) * total + i adds the table entries selected by the signed base-256 digits of
    // scalar (set, i) in windows [part * 32 / split, (part + 1) * 32 / split).  Digits: d_w = byte_w + carry, minus 256
    // (carry out) when that exceeds 128, so |d_w| <= 128; canonical scalars are < 2^255, the top byte is <= 0x40 and
    // the recoding never carries out of window 31.
    static H2_HD void accum_body(const FbPlan &p, const fe *scalars, const affine *dtab, xyzz *partial, uint64_t u) {
        if (u >= p.total * p.sets * p.split) return;
        const uint64_t i = u % p.total, sp = u / p.total;               // the lanes of a warp share (set, part): same windows,
        const uint32_t part = (uint32_t)(sp % p.split);                 // no divergence around the additions (with the part
        const uint64_t v = (sp / p.split) * p.total + i;                // innermost every warp ran 8 x 4 of them, measured)
        fe s = fe_load(scalars + v);
        if (p.scalars_mont) s = fe_from_mont<PS>(s);
        const uint32_t per = H2_FB_WINDOWS / p.split, w_lo = part * per, w_hi = w_lo + per;
        xyzz acc = xyzz_identity();
        uint32_t carry = 0;
        for (uint32_t w = 0; w < w_hi; w++) {
            uint32_t d = ((s.v[w >> 2] >> (8 * (w & 3))) & 0xffu) + carry;
            carry = d > 128u ? 1u : 0u;
            if (w < w_lo || d == 0 || d == 256u) continue;              // 256 - 256 = 0
            const uint32_t m = carry ? 256u - d : d;                    // |digit| in 1 .. 128
            affine pt = ld_affine(dtab + (i * H2_FB_WINDOWS + w) * H2_FB_MULTIPLES + (m - 1));
            if (carry) pt.y = fe_neg<P>(pt.y);
            xyzz_add_mixed<P>(acc, pt);
        }
        st_xyzz(partial + u, acc);
    }

    // Reduce, phase 1: quad q of CTA `cta` adds in[cta * 64 f + j * 64 + q], j < f (entries past `count` are the identity)
    static H2_HD xyzz reduce_gather(const xyzz *in, uint64_t count, uint32_t f, uint64_t cta, uint32_t q) {
        xyzz acc = xyzz_identity(), unused;
        for (uint32_t j = 0; j < f; j++) {
            const uint64_t idx = (cta * f + j) * H2_FB_QUADS + q;
            xyzz v = idx < count ? ld_xyzz(in + idx) : xyzz_identity(), r;
            xyzz_addsub_q<P, false>(acc, v, r, unused);
            acc = r;
        }
        return acc;
    }
    // phase 2, one tree level: acc += other (every quad of the CTA runs it: uniform control flow)
    static H2_HD xyzz reduce_level(const xyzz &acc, const xyzz &other) {
        xyzz r, unused;
        xyzz_addsub_q<P, false>(acc, other, r, unused);
        return r;
    }
    static H2_HD void finish(jacobian *out, const xyzz &total, uint32_t out_canonical) {
        jacobian j = xyzz_to_jacobian<P>(total);
        if (out_canonical) { j.x = fe_from_mont<P>(j.x); j.y = fe_from_mont<P>(j.y); j.z = fe_from_mont<P>(j.z); }
        st_jacobian(out, j);
    }
};

// host-side level plan of the reduce: `count` partial sums per set -> ceil(count / (64 f)) per level until one is left
inline uint32_t fb_fan(uint64_t count) {
    uint64_t f = (count + H2_FB_QUADS - 1) / H2_FB_QUADS;
    return (uint32_t)(f < 1 ? 1 : f > H2_FB_MAX_FAN ? H2_FB_MAX_FAN : f);
}
inline uint64_t fb_ctas(uint64_t count, uint32_t f) { return (count + (uint64_t)H2_FB_QUADS * f - 1) / ((uint64_t)H2_FB_QUADS * f); }
// threads per scalar: enough threads to fill the machine (>= 64 k) without shredding the per-thread chains -- and at most
// ~128 CTAs of partial sums for the first reduce level (one wave: its CTAs hold a whole SM's registers)
inline uint32_t fb_split(uint64_t total, uint32_t sets) {
    uint32_t split = 1;
    while (split < 8 && total * sets * split < (1ull << 16)) split <<= 1;
    return split;
}

#if defined(__CUDACC__)
template <class P, class PS> __global__ void __launch_bounds__(128) fb_table_kernel(const affine *wtab, affine *dtab, uint64_t count, uint64_t stride) {
    FixedBase<P, PS>::table_body(wtab, dtab, count, stride, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
template <class P, class PS> __global__ void __launch_bounds__(128, 5) fb_accum_kernel(const FbPlan p, const fe *scalars, const affine *dtab, xyzz *partial) {
    FixedBase<P, PS>::accum_body(p, scalars, dtab, partial, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
// grid (ctas, sets); in / out hold `in_stride` / `out_stride` entries per set.  final: the single CTA of a set writes
// the Jacobian result instead of an XYZZ partial.
template <class P, class PS>
__global__ void __launch_bounds__(4 * H2_FB_QUADS) fb_reduce_kernel(const xyzz *in, uint64_t count, uint64_t in_stride, uint32_t f, xyzz *out,
                                                                   uint64_t out_stride, jacobian *result, uint32_t out_canonical) {
    __shared__ xyzz sm[H2_FB_QUADS];
    const uint32_t q = threadIdx.x >> 2, set = blockIdx.y;
    xyzz acc = FixedBase<P, PS>::reduce_gather(in + (uint64_t)set * in_stride, count, f, blockIdx.x, q);
    st_xyzz_q(&sm[q], acc);
    for (uint32_t step = H2_FB_QUADS / 2; step >= 1; step >>= 1) {
        __syncthreads();
        xyzz other = q < step ? ld_xyzz(&sm[q + step]) : xyzz_identity();
        acc = FixedBase<P, PS>::reduce_level(acc, other);
        __syncthreads();
        if (q < step) st_xyzz_q(&sm[q], acc);
    }
    if (q == 0) {
        if (result) { if ((threadIdx.x & 3u) == 0) FixedBase<P, PS>::finish(result + set, acc, out_canonical); }
        else st_xyzz_q(out + (uint64_t)set * out_stride + blockIdx.x, acc);
    }
}
#endif

}  // namespace h2
