// Inner-product-argument round loop on the device (first "next" row).
//
// Replaces the body of the `for j in 0..k` loop of commitment::create_proof
// (halo2_proofs/src/poly/commitment/prover.rs:100-142): per round two half-size MSMs against the
// FOLDED generators g' (:107-108), two inner products (:110-111), the [value*z]U + [rand]W terms
// (:113-119), the scalar folds of p' and b (:134-139) and parallel_generator_collapse (:140, :154-166).
//
// B200 formulation: the generators are never folded.  After rounds 0..j-1 with challenges u_0..u_{j-1}
//     g'_i = sum_{t = i mod 2^(k-j)} s_t g_t,   s_t = prod_{j' < j, bit_{k-1-j'}(t) = 1} u_j'
// so with half = 2^(k-1-j) and bit = k-1-j
//     L_j = sum_{t: bit(t) = 0} (p'[(t mod half) + half] s_t) g_t + [value_l z] U + [l_rand] W
//     R_j = sum_{t: bit(t) = 1} (p'[ t mod half        ] s_t) g_t + [value_r z] U + [r_rand] W
// are two FIXED-BASE MSMs over the resident window table of g || w || u (msm.cuh `sets` = 2): one pass,
// no variable-base scalar multiplications and no batch normalisation per round.  Only p', b and the
// n coefficients s_t are folded (field work).  Results are the same group elements the reference
// computes; their affine encodings are therefore identical.
//
// All scalars live in Montgomery form in the SCALAR field PS.
#pragma once
#include "field.cuh"

namespace h2 {

struct IpaState {
    fe *p;        // p'      (n, first `2 * half` live)
    fe *b;        // b       (n, first `2 * half` live)
    fe *s;        // s_t     (n)
    fe *scal;     // 2 x (n + 2): [A_t | l_rand | value_l z] [B_t | r_rand | value_r z]
    uint64_t n;
};

template <class PS> struct Ipa {
    // A_t / B_t of round `bit` (half = 2^bit)
    static H2_HD void prep_body(const IpaState &S, uint32_t bit, uint64_t t) {
        if (t >= S.n) return;
        const uint64_t half = 1ull << bit, lo = t & (half - 1);
        const bool hi = (t >> bit) & 1u;
        fe st = fe_load(S.s + t);
        fe v = fe_mul<PS>(fe_load(S.p + (hi ? lo : lo + half)), st);
        fe_store(S.scal + t, hi ? fe_zero() : v);
        fe_store(S.scal + (S.n + 2) + t, hi ? v : fe_zero());
    }
    // partial inner products of thread `tid` of `nthr`: (sum p'[i + half] b[i], sum p'[i] b[i + half])
    static H2_HD void inner_partial(const IpaState &S, uint32_t bit, uint32_t tid, uint32_t nthr, fe &vl, fe &vr) {
        const uint64_t half = 1ull << bit;
        vl = fe_zero(); vr = fe_zero();
        for (uint64_t i = tid; i < half; i += nthr) {
            fe plo = fe_load(S.p + i), phi = fe_load(S.p + i + half), blo = fe_load(S.b + i), bhi = fe_load(S.b + i + half);
            vl = fe_add<PS>(vl, fe_mul<PS>(phi, blo));
            vr = fe_add<PS>(vr, fe_mul<PS>(plo, bhi));
        }
    }
    static H2_HD void inner_finish(const IpaState &S, const fe &vl, const fe &vr, const fe &z, const fe &l_rand, const fe &r_rand) {
        fe *a = S.scal + S.n, *bq = S.scal + (S.n + 2) + S.n;
        fe_store(a, l_rand); fe_store(a + 1, fe_mul<PS>(vl, z));
        fe_store(bq, r_rand); fe_store(bq + 1, fe_mul<PS>(vr, z));
    }
    // fold with challenge u (prover.rs:134-139) and fold it into the coefficients s_t
    static H2_HD void fold_body(const IpaState &S, uint32_t bit, const fe &u, const fe &u_inv, uint64_t t) {
        if (t >= S.n) return;
        const uint64_t half = 1ull << bit;
        if ((t >> bit) & 1u) fe_store(S.s + t, fe_mul<PS>(fe_load(S.s + t), u));
        if (t < half) {
            fe_store(S.p + t, fe_add<PS>(fe_load(S.p + t), fe_mul<PS>(fe_load(S.p + t + half), u_inv)));
            fe_store(S.b + t, fe_add<PS>(fe_load(S.b + t), fe_mul<PS>(fe_load(S.b + t + half), u)));
        }
    }
    static H2_HD void init_body(const IpaState &S, int p_is_mont, uint64_t t) {
        if (t >= S.n) return;
        if (!p_is_mont) fe_store(S.p + t, fe_to_mont<PS>(fe_load(S.p + t)));
        fe_store(S.s + t, fe_one<PS>());
    }
};

#ifdef __CUDACC__
template <class PS> __global__ void ipa_init_kernel(IpaState S, int p_is_mont) {
    Ipa<PS>::init_body(S, p_is_mont, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
template <class PS> __global__ void ipa_prep_kernel(IpaState S, uint32_t bit) {
    Ipa<PS>::prep_body(S, bit, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
template <class PS> __global__ void ipa_fold_kernel(IpaState S, uint32_t bit, fe u, fe u_inv) {
    Ipa<PS>::fold_body(S, bit, u, u_inv, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
// one CTA: strided partial sums, shared-memory tree, thread 0 writes the four tail scalars
template <class PS> __global__ void __launch_bounds__(512) ipa_inner_kernel(IpaState S, uint32_t bit, fe z, fe l_rand, fe r_rand) {
    __shared__ fe sl[512], sr[512];
    fe vl, vr;
    Ipa<PS>::inner_partial(S, bit, threadIdx.x, blockDim.x, vl, vr);
    sl[threadIdx.x] = vl; sr[threadIdx.x] = vr;
    __syncthreads();
    for (uint32_t step = blockDim.x >> 1; step > 0; step >>= 1) {
        if (threadIdx.x < step) {
            sl[threadIdx.x] = fe_add<PS>(sl[threadIdx.x], sl[threadIdx.x + step]);
            sr[threadIdx.x] = fe_add<PS>(sr[threadIdx.x], sr[threadIdx.x + step]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) Ipa<PS>::inner_finish(S, sl[0], sr[0], z, l_rand, r_rand);
}
// c = p'[0] (and b[0]) after the last fold, in the caller's representation
template <class PS> __global__ void ipa_result_kernel(IpaState S, int canonical, fe *out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        fe c = fe_load(S.p), b0 = fe_load(S.b);
        if (canonical) { c = fe_from_mont<PS>(c); b0 = fe_from_mont<PS>(b0); }
        fe_store(out, c); fe_store(out + 1, b0);
    }
}
#endif

}  // namespace h2
