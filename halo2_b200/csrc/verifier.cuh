// K18: the verifier's side of the path -- the g_scalars vector of the polynomial commitment scheme's MSM, resident.
//
//   compute_s            /root/reference/halo2_proofs/src/poly/commitment/verifier.rs:156-171
//                        s = the coefficients of g(X) = prod_{i<k} (1 + u_{k-1-i} X^(2^i)), times `init`; Guard::use_challenges
//                        (:36-41) adds compute_s(u, -c) to the MSM's g_scalars, Guard::compute_g (:58-62) commits to compute_s(u, 1)
//   MSM::add_to_g_scalars  poly/commitment/msm.rs:104-113     g_scalars[i] += scalars[i]
//   MSM::scale             poly/commitment/msm.rs:126-139     g_scalars[i] *= factor
//   MSM::add_msm           poly/commitment/msm.rs:37-62       the g_scalars part: ours[i] += theirs[i]
//
// The reference builds s by k doubling copies (v[len..2 len] = v[..len] * u_j), 2^k - 1 serial multiplies; element i is
// init * prod_{j : bit j of i set} u_{k-1-j}, independent of every other element, so a thread takes the four elements that
// share the bits above the lowest two: popcount(i >> 2) + 3 multiplies per four elements, and with `accumulate` the sum into
// g_scalars happens in the same pass (the vector s is never materialised).  BatchVerifier's `acc.scale(r); acc.add_msm(&msm)`
// (plonk/verifier/batch.rs:83-93) is one pass of scale_add.  Exact field arithmetic: THE elements the reference computes.
#pragma once
#include "field.cuh"

namespace h2 {

template <class P> struct VerifierOps {
    // elements [t << low, (t + 1) << low) of compute_s(u, init), low = min(k, 2); u: k challenges in Montgomery form, u[0] = u_0
    static H2_HD void compute_s_body(fe *dst, const fe *u, uint32_t k, const fe &init, int accumulate, uint64_t t) {
        const uint32_t low = k < 2 ? k : 2;
        if (t >= (1ull << (k - low))) return;
        fe s[4];
        s[0] = init;
        for (uint32_t j = low; j < k; j++)                       // bit j of the element index selects u_{k-1-j}
            if ((t >> (j - low)) & 1) s[0] = fe_mul<P>(s[0], fe_load(u + (k - 1 - j)));
        if (low >= 1) s[1] = fe_mul<P>(s[0], fe_load(u + (k - 1)));
        if (low == 2) {
            const fe u1 = fe_load(u + (k - 2));
            s[2] = fe_mul<P>(s[0], u1);
            s[3] = fe_mul<P>(s[1], u1);
        }
        fe *d = dst + (t << low);
        for (uint32_t e = 0; e < (1u << low); e++) fe_store(d + e, accumulate ? fe_add<P>(fe_load(d + e), s[e]) : s[e]);
    }
    // dst[i] = a * dst[i] + b * src[i]   (src == nullptr: dst[i] = a * dst[i])
    static H2_HD void scale_add_body(fe *dst, const fe *src, const fe &a, const fe &b, uint64_t n, uint64_t i) {
        if (i >= n) return;
        fe r = fe_mul<P>(fe_load(dst + i), a);
        if (src) r = fe_add<P>(r, fe_mul<P>(fe_load(src + i), b));
        fe_store(dst + i, r);
    }
};

#if defined(__CUDACC__)
template <class P> __global__ void __launch_bounds__(128) verifier_compute_s_kernel(fe *dst, const fe *u, uint32_t k, fe init, int accumulate) {
    VerifierOps<P>::compute_s_body(dst, u, k, init, accumulate, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
template <class P> __global__ void __launch_bounds__(256) verifier_scale_add_kernel(fe *dst, const fe *src, fe a, fe b, uint64_t n) {
    VerifierOps<P>::scale_add_body(dst, src, a, b, n, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
#endif

}  // namespace h2
