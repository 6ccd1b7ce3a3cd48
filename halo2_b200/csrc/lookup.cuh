// The lookup argument's permuted columns on resident polynomials.
//
// Replaces permute_expression_pair (/root/reference/halo2_proofs/src/plonk/lookup/prover.rs:563-647), the one step of the
// prover's middle section that is neither an FFT, an MSM nor an elementwise program: given the compressed input column A and
// table column S over the usable rows it returns
//   A' = A sorted (ff's Ord: the canonical integers, :577-581), and
//   S' with S'[r] = A'[r] on the first row of every run of equal values in A' (:595-603; the value must occur in S, else
//      Error::ConstraintSystemFailure, :605-608), the other rows filled with the table values that are left over, smallest
//      first, handed to the repeated rows from the LAST one down (`repeated_input_rows.pop()`, :617-622).
// The reference does this with a sort and a BTreeMap on one core.  Here: two bitonic sorts of 256-bit canonical keys (shared
// memory below 1024 keys, one launch per global stage above), a lower-bound search per first row that marks the table value it
// consumes, two exclusive scans (first rows, unconsumed table values) and two scatter / gather kernels.  The blinding rows
// (:625-627, random) stay with the caller.  Same values as the reference, position by position.
#pragma once
#include "field.cuh"

namespace h2 {

// lexicographic order of canonical (non-Montgomery) elements = order of the integers
H2_HD bool fe_canon_lt(const fe &a, const fe &b) {
    for (int i = 7; i >= 0; i--) {
        if (a.v[i] != b.v[i]) return a.v[i] < b.v[i];
    }
    return false;
}

#define H2_LK_BLOCK_LOG 10u           // keys per shared-memory block of the bitonic sort

template <class P> struct LookupPermute {
    // keys[i] = canonical src[i] for i < u, the all-ones sentinel (> every field element) up to the power of two N
    static H2_HD void load_body(const fe *src, uint64_t u, fe *keys, uint64_t N, uint64_t i) {
        if (i >= N) return;
        fe x;
        if (i < u) x = fe_from_mont<P>(fe_load(src + i));
        else for (int k = 0; k < 8; k++) x.v[k] = 0xffffffffu;
        fe_store(keys + i, x);
    }
    // one compare-exchange of the bitonic network: pair (i, i + stride) of the merge of `size` keys that i lies in
    static H2_HD void cex(fe &a, fe &b, bool ascending) {
        if (fe_canon_lt(b, a) == ascending) { fe t = a; a = b; b = t; }
    }
    static H2_HD void global_stage_body(fe *keys, uint64_t N, uint64_t size, uint64_t stride, uint64_t t) {
        if (t >= N / 2) return;
        const uint64_t i = (t / stride) * 2 * stride + (t % stride), j = i + stride;
        fe a = fe_load(keys + i), b = fe_load(keys + j);
        const bool before = fe_canon_lt(b, a);
        if (before == ((i & size) == 0)) { fe_store(keys + i, b); fe_store(keys + j, a); }
    }
    // rows r < u of the sorted input: out_input[r] = A'[r]; on the first row of a run also out_table[r] = A'[r] and the
    // table value it consumes is marked (lower bound in the sorted table; a miss raises *err).  first[r] = 1 / 0.
    static H2_HD void first_body(const fe *ka, const fe *kt, uint64_t u, uint32_t *first, uint32_t *unconsumed, uint32_t *err, fe *out_input,
                                 fe *out_table, uint64_t r) {
        if (r >= u) return;
        const fe v = fe_load(ka + r);
        fe_store(out_input + r, fe_to_mont<P>(v));
        bool is_first = r == 0;
        if (!is_first) is_first = !fe_eq(v, fe_load(ka + r - 1));
        first[r] = is_first ? 1u : 0u;
        if (!is_first) return;
        fe_store(out_table + r, fe_to_mont<P>(v));
        uint64_t lo = 0, hi = u;                                   // lower bound of v in kt[0, u)
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (fe_canon_lt(fe_load(kt + mid), v)) lo = mid + 1; else hi = mid;
        }
        if (lo >= u || !fe_eq(fe_load(kt + lo), v)) { *err = 1u; return; }
        unconsumed[lo] = 0u;                                       // distinct values have distinct lower bounds: no race
    }
    // leftover[rank] = the rank-th unconsumed table value (ranks from the exclusive scan of `unconsumed`)
    static H2_HD void leftover_body(const fe *kt, uint64_t u, const uint32_t *unconsumed_flag, const uint32_t *rank, fe *leftover, uint64_t i) {
        if (i < u && unconsumed_flag[i]) fe_store(leftover + rank[i], fe_load(kt + i));
    }
    // repeated rows take the leftovers from the back: the j-th repeated row (j = r - firsts_before[r]) gets leftover[L - 1 - j]
    static H2_HD void fill_body(uint64_t u, const uint32_t *first_flag, const uint32_t *firsts_before, const fe *leftover, fe *out_table, uint64_t r) {
        if (r >= u || first_flag[r]) return;
        const uint64_t total_first = firsts_before[u], L = u - total_first, j = r - firsts_before[r];
        fe_store(out_table + r, fe_to_mont<P>(fe_load(leftover + (L - 1 - j))));
    }
};

#if defined(__CUDACC__)
template <class P> __global__ void __launch_bounds__(256) lk_load_kernel(const fe *src, uint64_t u, fe *keys, uint64_t N) {
    LookupPermute<P>::load_body(src, u, keys, N, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
// Shared-memory part of the bitonic sort on a block of 2^H2_LK_BLOCK_LOG keys (or all N of them when N is smaller):
// size_lo == 2: every merge size 2 .. block (the block comes out sorted, direction by its global position);
// otherwise: the strides below the block size of the one merge of `size_lo` keys.
static __global__ void __launch_bounds__(512) lk_bitonic_block_kernel(fe *keys, uint64_t N, uint64_t size_lo, uint32_t full) {
    extern __shared__ uint4 lk_sm[];
    fe *sh = reinterpret_cast<fe *>(lk_sm);
    const uint32_t BL = (uint32_t)(N < (1ull << H2_LK_BLOCK_LOG) ? N : (1ull << H2_LK_BLOCK_LOG));
    const uint64_t base = (uint64_t)blockIdx.x * BL;
    for (uint32_t e = threadIdx.x; e < BL; e += blockDim.x) sh[e] = fe_load(keys + base + e);
    __syncthreads();
    for (uint64_t size = full ? 2 : size_lo; size <= (full ? BL : size_lo); size <<= 1) {
        for (uint32_t stride = (uint32_t)(size / 2 < BL / 2 ? size / 2 : BL / 2); stride >= 1; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < BL / 2; t += blockDim.x) {
                const uint32_t i = (t / stride) * 2 * stride + (t % stride), j = i + stride;
                fe a = sh[i], b = sh[j];
                const bool asc = ((base + i) & size) == 0;
                bool lt = false;
#pragma unroll
                for (int k = 7; k >= 0; k--) if (a.v[k] != b.v[k]) { lt = b.v[k] < a.v[k]; break; }
                if (lt == asc) { sh[i] = b; sh[j] = a; }
            }
            __syncthreads();
        }
    }
    for (uint32_t e = threadIdx.x; e < BL; e += blockDim.x) fe_store(keys + base + e, sh[e]);
}
template <class P> __global__ void __launch_bounds__(256) lk_bitonic_global_kernel(fe *keys, uint64_t N, uint64_t size, uint64_t stride) {
    LookupPermute<P>::global_stage_body(keys, N, size, stride, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
template <class P> __global__ void __launch_bounds__(128) lk_first_kernel(const fe *ka, const fe *kt, uint64_t u, uint32_t *first, uint32_t *unconsumed,
                                                                          uint32_t *err, fe *out_input, fe *out_table) {
    LookupPermute<P>::first_body(ka, kt, u, first, unconsumed, err, out_input, out_table, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
template <class P> __global__ void __launch_bounds__(256) lk_leftover_kernel(const fe *kt, uint64_t u, const uint32_t *flag, const uint32_t *rank, fe *leftover) {
    LookupPermute<P>::leftover_body(kt, u, flag, rank, leftover, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
template <class P> __global__ void __launch_bounds__(256) lk_fill_kernel(uint64_t u, const uint32_t *first_flag, const uint32_t *firsts_before,
                                                                         const fe *leftover, fe *out_table) {
    LookupPermute<P>::fill_body(u, first_flag, firsts_before, leftover, out_table, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
static __global__ void lk_fill_u32_kernel(uint32_t *a, uint64_t n, uint32_t v) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}
#endif

}  // namespace h2
