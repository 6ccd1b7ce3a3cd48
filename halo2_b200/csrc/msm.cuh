// Pippenger MSM kernels (K2..K5 in SURVEY.md section 2.1) for Pallas / Vesta.
//
// Replaces best_multiexp (/root/reference/halo2_proofs/src/arithmetic.rs:143-180) and its
// Buckets::sum inner loop (:74-93).  The reference walks all n points once per window on one
// core; here every (point, window) pair is an independent unit of work:
//
//   K2  digits + histogram   canonical scalar -> signed c-bit digits (zero digits skipped,
//                            like arithmetic.rs:78); count per (window, |digit|) bucket
//   K3  scan + scatter       counting sort of point references by global bucket id
//   K4  accumulate           fixed-size chunks of the sorted reference list, one chunk per
//                            thread, mixed XYZZ adds; a chunk is cut at bucket boundaries;
//                            pieces that do not cover a whole bucket become "partials" that
//                            the next (16x smaller) level merges.  Work per thread is
//                            constant, so skewed scalars (0/1 selector columns, all-equal
//                            scalars: SURVEY.md section 7 "MSM load imbalance") cost the same
//                            as uniform ones.
//   K5  bucket reduce        sum_b b * B[b] per window by hierarchical running sums (the
//                            parallel form of arithmetic.rs:86-92), then the window combine
//                            sum_w 2^(c w) S_w (arithmetic.rs:163,166).
//
// Signed digits halve the bucket count: digit d in [-(2^(c-1) - 1), 2^(c-1)], a negative digit
// adds the negated base to bucket |d|.  The group element computed is identical to the
// reference's; the ABI returns it as a Jacobian point and the tests compare affine canonical
// bytes.
#pragma once
#include "curve.cuh"

namespace h2 {

H2_HD affine ld_affine(const affine *p) { affine r; r.x = fe_load(&p->x); r.y = fe_load(&p->y); return r; }
H2_HD void st_affine(affine *p, const affine &a) { fe_store(&p->x, a.x); fe_store(&p->y, a.y); }
H2_HD xyzz ld_xyzz(const xyzz *p) { xyzz r; r.x = fe_load(&p->x); r.y = fe_load(&p->y); r.zz = fe_load(&p->zz); r.zzz = fe_load(&p->zzz); return r; }
H2_HD void st_xyzz(xyzz *p, const xyzz &a) { fe_store(&p->x, a.x); fe_store(&p->y, a.y); fe_store(&p->zz, a.zz); fe_store(&p->zzz, a.zzz); }
H2_HD void st_jacobian(jacobian *p, const jacobian &a) { fe_store(&p->x, a.x); fe_store(&p->y, a.y); fe_store(&p->z, a.z); }

#define H2_MSM_INVALID_KEY 0xffffffffu
#define H2_MSM_MAX_LEVELS 12

struct MsmPlan {
    uint64_t n;          // number of (scalar, base) pairs
    uint32_t c;          // window bits
    uint32_t W;          // windows = ceil(256 / c)
    uint32_t B;          // buckets per window = 2^(c-1)
    uint64_t G;          // total buckets = W * B
    uint64_t max_refs;   // n * W
    // accumulate levels: level 0 consumes refs, level i>0 consumes partial slots
    uint32_t acc_levels;
    uint32_t acc_chunk[H2_MSM_MAX_LEVELS];     // items per thread
    uint64_t acc_threads[H2_MSM_MAX_LEVELS];   // threads at that level
    uint64_t acc_slots[H2_MSM_MAX_LEVELS];     // INPUT slots of level i (level 0: max_refs)
    uint64_t part_offset[H2_MSM_MAX_LEVELS];   // offset of level-i input slots in the partial arrays (i >= 1)
    uint64_t part_total;                       // total partial slots
    // bucket-reduce levels
    uint32_t red_levels;
    uint32_t red_log_l[H2_MSM_MAX_LEVELS];     // log2 chunk length of level i
    uint32_t red_m_in[H2_MSM_MAX_LEVELS];      // entries per window entering level i
    uint32_t red_dbl[H2_MSM_MAX_LEVELS];       // doublings applied to acc at level i
    uint64_t red_offset[H2_MSM_MAX_LEVELS];    // offset of level-i OUTPUT in sums/E arrays
    uint64_t red_total;
};

inline uint32_t msm_default_window(uint64_t n) {
    // Tuned for XYZZ cost model: n*W mixed adds (10 M) + 2*W*2^(c-1) full adds (14 M).
    if (n < 32) return 3;
    uint32_t lg = 0;
    while ((1ull << (lg + 1)) <= n) lg++;
    uint32_t c = lg > 4 ? lg - 4 : 1;
    if (c < 4) c = 4;
    if (c > 20) c = 20;
    return c;
}

inline void msm_make_plan(MsmPlan &p, uint64_t n, uint32_t c) {
    p.n = n; p.c = c;
    p.W = (256 + c - 1) / c;
    p.B = 1u << (c - 1);
    p.G = (uint64_t)p.W * p.B;
    p.max_refs = n * p.W;
    // level 0 chunk: aim for >= 64K threads, between 4 and 32 refs each
    uint64_t k0 = p.max_refs / 65536;
    if (k0 < 4) k0 = 4;
    if (k0 > 32) k0 = 32;
    uint32_t lv = 0;
    uint64_t slots = p.max_refs;
    p.part_total = 0;
    p.part_offset[0] = 0;            // level 0 reads refs/keys, not partial slots
    for (;;) {
        uint32_t chunk = lv == 0 ? (uint32_t)k0 : 16u;
        uint64_t threads = (slots + chunk - 1) / chunk;
        if (threads == 0) threads = 1;
        p.acc_chunk[lv] = chunk; p.acc_threads[lv] = threads; p.acc_slots[lv] = slots;
        lv++;
        if (threads == 1) break;     // a single thread sees every remaining piece: all flushes complete
        // the 2 output slots per thread of this level are the input of the next
        p.part_offset[lv] = p.part_total;
        slots = 2 * threads;
        p.part_total += slots;
    }
    p.acc_levels = lv;
    // bucket reduce: level 0 chunks of 8 (parallel), deeper levels of 4 (short serial chains)
    uint32_t m = p.B, rl = 0, dbl = 0;
    p.red_total = 0;
    while (true) {
        uint32_t l = rl == 0 ? 3u : 2u;
        p.red_log_l[rl] = l; p.red_m_in[rl] = m; p.red_dbl[rl] = dbl;
        uint32_t m_out = (m + (1u << l) - 1) >> l;
        p.red_offset[rl] = p.red_total;
        p.red_total += (uint64_t)p.W * m_out;
        dbl += l;
        rl++;
        m = m_out;
        if (m == 1) break;
    }
    p.red_levels = rl;
}

struct MsmBuffers {
    // inputs
    const fe *scalars;        // n, canonical or Montgomery (see scalars_mont)
    const affine *bases;      // n, Montgomery coordinates
    uint32_t scalars_mont;
    // scratch
    fe *scal_canon;           // n (only when scalars_mont)
    uint32_t *counts;         // G + 1  (histogram, then exclusive offsets after the scan)
    uint32_t *cursor;         // G
    uint32_t *refs;           // max_refs   point index | sign << 31
    uint32_t *keys;           // max_refs   global bucket id
    xyzz *bucket_sum;         // G
    uint32_t *pkey, *pstart, *pend;   // part_total
    xyzz *ppt;                // part_total
    xyzz *red_sums, *red_e;   // red_total
    xyzz *win_sums;           // W
    jacobian *result;         // 1
};

// ---------------------------------------------------------------------------------------------
template <class P, class PS> struct Msm {
    // signed window digit; `carry` threads through the windows in increasing order
    static H2_HD int32_t next_digit(const uint32_t (&s)[8], uint32_t w, uint32_t c, uint32_t &carry) {
        uint32_t bit = w * c, idx = bit >> 5, sh = bit & 31;
        uint64_t v = 0;
        // register-resident select instead of dynamic indexing
        for (int i = 0; i < 8; i++) {
            if ((uint32_t)i == idx) v |= s[i];
            if ((uint32_t)i == idx + 1) v |= (uint64_t)s[i] << 32;
        }
        uint32_t raw = (uint32_t)(v >> sh) & ((1u << c) - 1u);
        uint32_t d = raw + carry;
        if (d > (1u << (c - 1))) { carry = 1; return (int32_t)d - (int32_t)(1u << c); }
        carry = 0;
        return (int32_t)d;
    }

    static H2_HD void load_scalar(const MsmBuffers &M, uint64_t i, uint32_t (&s)[8], bool first_pass) {
        fe x;
        if (M.scalars_mont) {
            if (first_pass) { x = fe_from_mont<PS>(fe_load(M.scalars + i)); fe_store(M.scal_canon + i, x); }
            else x = fe_load(M.scal_canon + i);
        } else x = fe_load(M.scalars + i);
        for (int k = 0; k < 8; k++) s[k] = x.v[k];
    }

    // ---- K4 helpers
    struct Flusher {
        const MsmPlan *plan; const MsmBuffers *M;
        uint64_t out_base;     // first output slot of this thread (2 per thread), or ~0 if last level
        uint32_t used;
        H2_HD void flush(uint32_t g, uint32_t a, uint32_t b, const xyzz &acc) {
            uint32_t lo = M->counts[g], hi = M->counts[g + 1];
            if (a == lo && b == hi) {
                st_xyzz(M->bucket_sum + g, acc);
            } else {
                uint64_t slot = out_base + used;
                used++;
                M->pkey[slot] = g; M->pstart[slot] = a; M->pend[slot] = b;
                st_xyzz(M->ppt + slot, acc);
            }
        }
    };

    // level 0: chunk of the sorted reference list
    static H2_HD void accum0_body(const MsmPlan &p, const MsmBuffers &M, uint64_t t) {
        const uint64_t total = M.counts[p.G];
        const uint32_t K = p.acc_chunk[0];
        uint64_t start = t * K;
        if (start >= total) return;
        uint64_t end = start + K < total ? start + K : total;
        Flusher F; F.plan = &p; F.M = &M; F.used = 0;
        F.out_base = p.acc_levels > 1 ? p.part_offset[1] + 2 * t : 0;
        xyzz acc = xyzz_identity();
        uint32_t cur = M.keys[start];
        uint32_t seg = (uint32_t)start;
        for (uint64_t pos = start; pos < end; pos++) {
            uint32_t g = M.keys[pos];
            if (g != cur) {
                F.flush(cur, seg, (uint32_t)pos, acc);
                acc = xyzz_identity(); cur = g; seg = (uint32_t)pos;
            }
            uint32_t ref = M.refs[pos];
            affine b = ld_affine(M.bases + (ref & 0x7fffffffu));
            if (ref >> 31) b.y = fe_neg<P>(b.y);
            xyzz_add_mixed<P>(acc, b);
        }
        F.flush(cur, seg, (uint32_t)end, acc);
    }

    // level >= 1: chunk of partial slots
    static H2_HD void accumN_body(const MsmPlan &p, const MsmBuffers &M, uint32_t lv, uint64_t t) {
        const uint32_t K = p.acc_chunk[lv];
        const uint64_t in_base = p.part_offset[lv], slots = p.acc_slots[lv];
        uint64_t start = t * K;
        if (start >= slots) return;
        uint64_t end = start + K < slots ? start + K : slots;
        Flusher F; F.plan = &p; F.M = &M; F.used = 0;
        F.out_base = lv + 1 < p.acc_levels ? p.part_offset[lv + 1] + 2 * t : 0;
        bool have = false;
        xyzz acc = xyzz_identity();
        uint32_t cur = 0, a = 0, b = 0;
        for (uint64_t s = in_base + start; s < in_base + end; s++) {
            uint32_t g = M.pkey[s];
            if (g == H2_MSM_INVALID_KEY) continue;
            if (have && g == cur) {
                xyzz_add<P>(acc, ld_xyzz(M.ppt + s));
                b = M.pend[s];
            } else {
                if (have) F.flush(cur, a, b, acc);
                have = true; cur = g; a = M.pstart[s]; b = M.pend[s]; acc = ld_xyzz(M.ppt + s);
            }
        }
        if (have) F.flush(cur, a, b, acc);
    }

    // ---- K5: one level of the hierarchical weighted sum.  Thread u of window w reduces
    // entries [u L, u L + L) of its input.
    static H2_HD void reduce_body(const MsmPlan &p, const MsmBuffers &M, uint32_t lv, uint64_t tid) {
        const uint32_t l = p.red_log_l[lv], L = 1u << l, m = p.red_m_in[lv];
        const uint32_t m_out = (m + L - 1) >> l;
        if (tid >= (uint64_t)p.W * m_out) return;
        uint32_t w = (uint32_t)(tid / m_out), u = (uint32_t)(tid % m_out);
        const xyzz *A = lv == 0 ? M.bucket_sum + (uint64_t)w * p.B
                                : M.red_sums + p.red_offset[lv - 1] + (uint64_t)w * m;
        uint32_t lo = u * L, hi = lo + L < m ? lo + L : m;
        xyzz run = xyzz_identity(), acc = xyzz_identity();
        for (uint32_t i = hi - 1; i > lo; i--) {
            xyzz_add<P>(run, ld_xyzz(A + i));
            xyzz_add<P>(acc, run);
        }
        xyzz_add<P>(run, ld_xyzz(A + lo));
        for (uint32_t d = 0; d < p.red_dbl[lv]; d++) xyzz_double<P>(acc);
        if (lv > 0) {
            const xyzz *E = M.red_e + p.red_offset[lv - 1] + (uint64_t)w * m;
            for (uint32_t i = lo; i < hi; i++) xyzz_add<P>(acc, ld_xyzz(E + i));
        }
        uint64_t o = p.red_offset[lv] + (uint64_t)w * m_out + u;
        st_xyzz(M.red_sums + o, run);
        st_xyzz(M.red_e + o, acc);
    }

    // window sum S_w = F + total (bucket weights are idx + 1), then 2^(c w) S_w
    static H2_HD xyzz window_value(const MsmPlan &p, const MsmBuffers &M, uint32_t w) {
        uint64_t o = p.red_offset[p.red_levels - 1] + w;
        xyzz s = ld_xyzz(M.red_e + o);
        xyzz_add<P>(s, ld_xyzz(M.red_sums + o));
        for (uint32_t d = 0; d < p.c * w; d++) xyzz_double<P>(s);
        return s;
    }
    static H2_HD void finish(const MsmBuffers &M, const xyzz &total, uint32_t out_canonical) {
        jacobian j = xyzz_to_jacobian<P>(total);
        if (out_canonical) { j.x = fe_from_mont<P>(j.x); j.y = fe_from_mont<P>(j.y); j.z = fe_from_mont<P>(j.z); }
        st_jacobian(M.result, j);
    }
};

#if defined(__CUDACC__)
// Warp-aggregated "fetch and add 1" on per-lane addresses: lanes that hit the same counter
// elect a leader that does one atomicAdd for the group (all-equal scalars would otherwise
// serialise n atomics on one L2 address).
__device__ __forceinline__ uint32_t agg_inc(uint32_t *ctr, bool active) {
    unsigned mask = __ballot_sync(0xffffffffu, active);
    if (!active) return 0;
    unsigned peers = __match_any_sync(mask, (unsigned long long)ctr);
    unsigned lane = threadIdx.x & 31u;
    int leader = __ffs(peers) - 1;
    uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(ctr, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    return base + rank;
}

template <class P, class PS> __global__ void __launch_bounds__(256) msm_hist_kernel(const MsmPlan p, const MsmBuffers M) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool in = i < p.n;
    uint32_t s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (in) Msm<P, PS>::load_scalar(M, i, s, true);
    uint32_t carry = 0;
    for (uint32_t w = 0; w < p.W; w++) {
        int32_t d = Msm<P, PS>::next_digit(s, w, p.c, carry);
        bool act = in && d != 0;
        uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
        uint64_t g = (uint64_t)w * p.B + (act ? mag - 1 : 0);
        agg_inc(M.counts + g, act);
    }
}
template <class P, class PS> __global__ void __launch_bounds__(256) msm_scatter_kernel(const MsmPlan p, const MsmBuffers M) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool in = i < p.n;
    uint32_t s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (in) Msm<P, PS>::load_scalar(M, i, s, false);
    uint32_t carry = 0;
    for (uint32_t w = 0; w < p.W; w++) {
        int32_t d = Msm<P, PS>::next_digit(s, w, p.c, carry);
        bool act = in && d != 0;
        uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
        uint64_t g = (uint64_t)w * p.B + (act ? mag - 1 : 0);
        uint32_t r = agg_inc(M.cursor + g, act);
        if (act) {
            uint32_t pos = M.counts[g] + r;
            M.refs[pos] = (uint32_t)i | (d < 0 ? 0x80000000u : 0u);
            M.keys[pos] = (uint32_t)g;
        }
    }
}
template <class P, class PS> __global__ void __launch_bounds__(128) msm_accum0_kernel(const MsmPlan p, const MsmBuffers M) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < p.acc_threads[0]) Msm<P, PS>::accum0_body(p, M, t);
}
template <class P, class PS> __global__ void __launch_bounds__(128) msm_accumN_kernel(const MsmPlan p, const MsmBuffers M, uint32_t lv) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < p.acc_threads[lv]) Msm<P, PS>::accumN_body(p, M, lv, t);
}
template <class P, class PS> __global__ void __launch_bounds__(128) msm_reduce_kernel(const MsmPlan p, const MsmBuffers M, uint32_t lv) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    Msm<P, PS>::reduce_body(p, M, lv, t);
}
// one CTA of 32 threads: thread w shifts its window, then a shared-memory tree adds them
template <class P, class PS> __global__ void __launch_bounds__(32) msm_combine_kernel(const MsmPlan p, const MsmBuffers M, uint32_t out_canonical) {
    __shared__ xyzz sh[32];
    uint32_t w = threadIdx.x;
    xyzz v = xyzz_identity();
    // W can exceed 32 for tiny windows: thread w takes windows w, w+32, ...
    for (uint32_t ww = w; ww < p.W; ww += 32) { xyzz t = Msm<P, PS>::window_value(p, M, ww); xyzz_add<P>(v, t); }
    sh[w] = v;
    __syncwarp();
    for (uint32_t off = 16; off > 0; off >>= 1) {
        if (w < off) { xyzz a = sh[w]; xyzz_add<P>(a, sh[w + off]); sh[w] = a; }
        __syncwarp();
    }
    if (w == 0) Msm<P, PS>::finish(M, sh[0], out_canonical);
}

// exclusive scan of counts[0..G] in place (counts[G] becomes the total), three small kernels
#define H2_SCAN_BLOCK 1024
#define H2_SCAN_ITEMS 8
__global__ void __launch_bounds__(H2_SCAN_BLOCK) scan_block_sums_kernel(const uint32_t *in, uint64_t n, uint32_t *block_sums) {
    __shared__ uint32_t sh[32];
    uint64_t base = (uint64_t)blockIdx.x * H2_SCAN_BLOCK * H2_SCAN_ITEMS;
    uint32_t s = 0;
    for (int k = 0; k < H2_SCAN_ITEMS; k++) {
        uint64_t i = base + (uint64_t)threadIdx.x * H2_SCAN_ITEMS + k;
        if (i < n) s += in[i];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = sh[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = s;
    }
}
__global__ void __launch_bounds__(H2_SCAN_BLOCK) scan_single_block_kernel(uint32_t *a, uint32_t n) {
    // exclusive scan of a[0..n) by one block, n arbitrary (loops in tiles of blockDim)
    __shared__ uint32_t sh[H2_SCAN_BLOCK];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += H2_SCAN_BLOCK) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < n ? a[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (uint32_t o = 1; o < H2_SCAN_BLOCK; o <<= 1) {
            uint32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        uint32_t incl = sh[threadIdx.x], c = carry_s;
        if (i < n) a[i] = c + incl - v;
        __syncthreads();
        if (threadIdx.x == H2_SCAN_BLOCK - 1) carry_s = c + incl;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(H2_SCAN_BLOCK) scan_apply_kernel(uint32_t *a, uint64_t n, const uint32_t *block_offsets) {
    __shared__ uint32_t sh[H2_SCAN_BLOCK];
    uint64_t base = (uint64_t)blockIdx.x * H2_SCAN_BLOCK * H2_SCAN_ITEMS + (uint64_t)threadIdx.x * H2_SCAN_ITEMS;
    uint32_t v[H2_SCAN_ITEMS], s = 0;
    for (int k = 0; k < H2_SCAN_ITEMS; k++) { v[k] = base + k < n ? a[base + k] : 0; s += v[k]; }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (uint32_t o = 1; o < H2_SCAN_BLOCK; o <<= 1) {
        uint32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    uint32_t run = block_offsets[blockIdx.x] + sh[threadIdx.x] - s;
    for (int k = 0; k < H2_SCAN_ITEMS; k++) {
        if (base + k < n) a[base + k] = run;
        run += v[k];
    }
}
#endif  // __CUDACC__

}  // namespace h2
