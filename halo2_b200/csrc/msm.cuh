// Pippenger MSM kernels (K2..K5 in SURVEY.md section 2.1) for Pallas / Vesta.
//
// Replaces best_multiexp (/root/reference/halo2_proofs/src/arithmetic.rs:143-180) and its
// Buckets::sum inner loop (:74-93).  The reference walks all n points once per window on one
// core; here every (point, window) pair is an independent unit of work:
//
//   K2  digits + histogram   canonical scalar -> signed c-bit digits (zero digits skipped,
//                            like arithmetic.rs:78); count per (window, |digit|) bucket
//   K3  scan + scatter       counting sort of point references by global bucket id
//   K4  accumulate           one work item per bucket (mixed XYZZ adds over its references), items
//                            ordered by size, largest first, so the lanes of a warp carry equal
//                            work and long items start early.  A bucket with more than T
//                            references is split into T-sized items whose results ("partials")
//                            are merged by 8x-shrinking levels: 0/1 selector columns, all-equal
//                            scalars etc. (SURVEY.md section 7 "MSM load imbalance") stay
//                            parallel, while ordinary inputs produce no partials at all.
//   K5  bucket reduce        S_w = sum_b b * B[b] per window (the parallel form of
//                            arithmetic.rs:86-92): chunked running sums (level A), then masked
//                            tree sums D_k = sum of chunk totals whose index has bit k set, so
//                            that S_w = E + T + 2^l0 * sum_k 2^k D_k needs one short Horner.
//       window combine       R = sum_w 2^(c w) S_w (arithmetic.rs:163,166): every (window, row)
//                            term shifted in parallel, then tree sums.
//
// Signed digits halve the bucket count: digit d in [-(2^(c-1) - 1), 2^(c-1)], a negative digit
// adds the negated base to bucket |d|.  The group element computed is identical to the
// reference's; the ABI returns it as a Jacobian point and the tests compare affine canonical
// bytes.
#pragma once
#include "curve.cuh"
#include "glv.cuh"

namespace h2 {

H2_HD affine ld_affine(const affine *p) { affine r; r.x = fe_load(&p->x); r.y = fe_load(&p->y); return r; }
H2_HD void st_affine(affine *p, const affine &a) { fe_store(&p->x, a.x); fe_store(&p->y, a.y); }
H2_HD xyzz ld_xyzz(const xyzz *p) { xyzz r; r.x = fe_load(&p->x); r.y = fe_load(&p->y); r.zz = fe_load(&p->zz); r.zzz = fe_load(&p->zzz); return r; }
H2_HD void st_xyzz(xyzz *p, const xyzz &a) { fe_store(&p->x, a.x); fe_store(&p->y, a.y); fe_store(&p->zz, a.zz); fe_store(&p->zzz, a.zzz); }
H2_HD void st_jacobian(jacobian *p, const jacobian &a) { fe_store(&p->x, a.x); fe_store(&p->y, a.y); fe_store(&p->z, a.z); }

#define H2_MSM_INVALID_KEY 0xffffffffu
#define H2_MSM_MAX_LEVELS 12
#define H2_R0_LOG 3           // R0 block = 8 level-A entries
#define H2_R0_ROWS 5          // per block: T, E, D_0..D_2

struct MsmPlan {
    uint64_t n;          // number of (scalar, base) pairs
    uint32_t c;          // window bits
    uint32_t glv;        // 1: scalars are split k = k1 + k2 lambda (glv.cuh); point i contributes P_i (k1) and phi(P_i) (k2)
    uint32_t chunks;     // bucket_sum holds `chunks` arrays of G partial bucket sums (one per upload chunk of the bases), summed by level A
    uint32_t cap;        // bin capacity of the single-pass sort (0: exact two-pass sort only)
    uint32_t cap_top;    // ... of the top window's bins (its digits are fewer bits wide and, with GLV, not uniform)
    uint64_t g_top;      // first bucket of the top window (= G in the fixed-base mode: no separate top region)
    uint32_t top_bins;   // populated buckets of the top window: digits 1 .. top_bins
    uint64_t ref_space;  // entries of the refs array
    uint32_t W;          // windows = ceil(256 / c), or ceil(128 / c) with the GLV split
    uint32_t B;          // buckets per window = 2^(c-1)
    uint32_t fixed;      // 1: bases come from a precomputed table T[w][i] = 2^(c w) G_i (resident Params
                         //    generators): every window's digits share ONE bucket set, no window combine
    uint32_t sets;       // fixed: independent MSMs over the same table in one pass (polynomial batch); scalars
                         //    are laid out [set][n], results [set]
    uint32_t Wb;         // bucket sets = fixed ? sets : W
    uint64_t stride;     // fixed: points per table window
    uint64_t G;          // total buckets = Wb * B
    uint64_t max_refs;   // n * W (x 2 with the GLV split)
    uint32_t T;          // references per work item (larger buckets are split)
    uint32_t natural;    // fast pass whose work items are simply the buckets in index order (no size sort: when every lane-group is resident at
                         // once the kernel lasts as long as the longest bucket whatever the order, and three item kernels go away)
    uint32_t fast;       // fixed-base pass WITHOUT the fallback kernels (exact sort, partial merges): valid only if the device flags stay
                         // clear -- the host checks them with the result and re-runs the full pass otherwise (capi_msm.cu)
    uint32_t ba;         // batched-affine rounds before the XYZZ chain (0: none), see Msm::ba_round_body
    uint32_t ba_m[4];    // work items per thread in round r = 1 .. ba (index r - 1)
    uint64_t max_items;  // upper bound of work items = G + max_refs / T
    // partial-merge levels: level 1 consumes the slots written by split buckets
    uint32_t acc_levels;                       // level 0 = the bucket items
    uint32_t acc_chunk[H2_MSM_MAX_LEVELS];     // slots per thread (levels >= 1)
    uint64_t acc_threads[H2_MSM_MAX_LEVELS];   // threads at that level (upper bound)
    uint64_t acc_slots[H2_MSM_MAX_LEVELS];     // INPUT slots of level i (i >= 1)
    uint64_t part_offset[H2_MSM_MAX_LEVELS];   // offset of level-i input slots in the partial arrays
    uint64_t part_total;
    // bucket reduce
    uint32_t l0;         // log2 of the level-A chunk
    uint32_t m1;         // entries per window after level A = B >> l0   (power of two)
    uint32_t nb0;        // R0 blocks (8 entries) per window = ceil(m1 / 8)   (power of two)
    uint32_t bits0;      // bits produced by R0 = min(3, log2 m1)
    uint32_t bits1;      // bits produced by R1 = log2 nb0
    uint32_t r1_rows;    // 2 + bits0 + bits1
};

inline uint32_t ilog2_u32(uint32_t v) { uint32_t r = 0; while (r < 31 && (1u << (r + 1)) <= v) r++; return r; }

// Window size.  Besides the usual work balance (n W mixed adds vs 2 W 2^(c-1) reduce adds) the TOP window
// matters: it only holds (bits mod c) real bits, so its references crowd into few buckets unless that
// remainder is close to c.  255 = 15*16 + 15 = 19*13 + 8 = 31*8 + 7 and (GLV halves) 127 = 7*16 + 15 =
// 9*13 + 10 = 15*8 + 7: c in {8, 13, 16} keeps the top window well filled for both scalar widths.
inline uint32_t msm_default_window(uint64_t n, uint32_t glv = 0) {
    uint64_t n_eff = glv ? 2 * n : n;
    if (n_eff < 64) return 4;
    if (n_eff < (1ull << 10)) return 8;
    if (n_eff < (1ull << 17)) return 13;     // measured: 13 beats 16 up to n = 2^15 (GLV), tools/sweep_window.py
    if (n_eff < (1ull << 25)) return 16;     // ... 16 wins from 2^16 to 2^22 (8 full windows with GLV)
    return 19;                               // ... 19 (7 windows) from 2^24
}

#define H2_MSM_NO_BINS 0xffffffffu
inline void msm_make_plan(MsmPlan &p, uint64_t n, uint32_t c, uint32_t force_t = 0, uint32_t force_kn = 0, uint32_t fixed = 0,
                          uint64_t stride = 0, uint32_t glv = 0, uint32_t sets = 1, uint32_t force_cap = 0) {
    p.n = n; p.c = c; p.chunks = 1;
    p.fast = 0; p.natural = 0;
    p.ba = 0; p.ba_m[0] = p.ba_m[1] = p.ba_m[2] = p.ba_m[3] = 1;
    p.glv = fixed ? 0u : glv;
    // GLV sub-scalars are < 2^127 (glv.cuh): with W c >= 128 the top window's raw digit is < 2^(c-1), so it
    // absorbs the signed-digit carry without opening another window
    p.W = p.glv ? (128 + c - 1) / c : (256 + c - 1) / c;
    p.B = 1u << (c - 1);
    p.sets = fixed && sets ? sets : 1u;
    p.fixed = fixed; p.Wb = fixed ? p.sets : p.W; p.stride = stride;
    p.G = (uint64_t)p.Wb * p.B;
    p.max_refs = n * p.W * (p.glv ? 2 : 1) * p.sets;
    // references per work item: 128 for big problems; shorter chains when there is little parallelism
    uint32_t T = 256;
    while (T > 32 && p.max_refs / T < 65536) T >>= 1;
    p.T = force_t ? force_t : T;
    p.max_items = p.G + p.max_refs / p.T + 1;
    // Single-pass sort: bucket g owns a bin of `cap` references (see Msm::bucket_lo).  With mean load lambda the capacity
    // lambda + 8 sqrt(lambda) + 16 is never reached by uniformly random digits; inputs that do overflow a bin
    // (repeated scalars, 0/1 columns) fall back to the exact histogram / scan / scatter sort (flags[1]).
    // The top window of a one-shot MSM is different: it holds top_bits < c bits, so its digits crowd into 2^top_bits
    // buckets, and GLV halves (alpha v1 + beta v2, alpha, beta uniform in [-1/2, 1/2)) have trapezoid densities peaking
    // near zero -- its low buckets receive 2^127 (1 / max(a1, a2) + 1 / max(|b1|, b2)) = 2.6 x the load of a uniform
    // 127-bit value.  It gets its own bin count and capacity.
    {
        auto capacity = [](uint64_t lam) { uint64_t rt = 0; while ((rt + 1) * (rt + 1) <= lam) rt++; return (lam + 8 * rt + 16 + 7) & ~7ull; };
        const uint64_t n_eff = (p.glv ? 2 : 1) * n;
        uint64_t cap, cap_top = 0, top_bins = 0;
        p.g_top = p.G;
        if (fixed || p.W < 2) {
            // all windows share one bucket set; the top window (254 - (W - 1) c bits) adds n >> top_bits references to
            // each of its low buckets.  (When that is far above the mean the bins overflow and the exact sort runs:
            // table_window() avoids such window sizes.)
            uint64_t lam = (p.max_refs + p.G - 1) / p.G;
            const int tb = 254 - (int)((p.W - 1) * c);
            if (fixed && tb > 0 && tb < (int)c - 1 && (n >> tb) <= 4 * lam + 16) lam += (n >> tb) + 1;
            cap = capacity(lam);
        } else {
            // scalars are < 2^254 (1 + 2^-129): 254 significant bits; GLV halves < 2^127
            const int bits = p.glv ? 127 : 254, tb = bits - (int)((p.W - 1) * c), top_bits = tb < 0 ? 0 : (tb > (int)c - 1 ? (int)c - 1 : tb);
            cap = capacity((n_eff + p.B - 1) / p.B);
            top_bins = (1ull << top_bits) + 1 < p.B ? (1ull << top_bits) + 1 : p.B;
            uint64_t lam_top = ((p.glv ? n * 53 / 10 : n) >> top_bits) + 1;
            cap_top = capacity(lam_top);
            if (cap_top > n_eff + 8) cap_top = (n_eff + 8) & ~7ull;
            p.g_top = (uint64_t)(p.W - 1) * p.B;
        }
        uint64_t space = p.g_top * cap + top_bins * cap_top;
        if (force_cap == H2_MSM_NO_BINS || space >= (1ull << 32)) { cap = 0; cap_top = 0; }
        else if (force_cap) { cap = force_cap; cap_top = cap_top ? force_cap : 0; }
        p.cap = (uint32_t)cap; p.cap_top = (uint32_t)cap_top; p.top_bins = (uint32_t)top_bins;
        space = p.g_top * p.cap + (uint64_t)p.top_bins * p.cap_top;
        p.ref_space = space > p.max_refs ? space : p.max_refs;
    }
    // partial slots: slot(start, chunk) = 2 * (start / T) + (chunk > 0), see item_slot()
    uint32_t lv = 1;
    uint64_t slots = 2 * (p.ref_space / p.T + 1);
    p.acc_chunk[0] = p.T; p.acc_threads[0] = p.max_items; p.acc_slots[0] = p.max_refs; p.part_offset[0] = 0;
    p.part_offset[1] = 0;
    p.part_total = slots;
    for (;;) {
        uint32_t chunk = force_kn ? force_kn : 16u;
        if (lv == H2_MSM_MAX_LEVELS - 1) chunk = (uint32_t)slots;   // last allowed level: one thread takes the rest
        uint64_t threads = (slots + chunk - 1) / chunk;
        if (threads == 0) threads = 1;
        p.acc_chunk[lv] = chunk; p.acc_threads[lv] = threads; p.acc_slots[lv] = slots;
        lv++;
        if (threads == 1) break;     // a single thread sees every remaining piece: all flushes complete
        p.part_offset[lv] = p.part_total;
        slots = 2 * threads;         // the 2 output slots per thread are the input of the next level
        p.part_total += slots;
    }
    p.acc_levels = lv;
    // bucket reduce
    // level-A chunk: 8 buckets when there is plenty of parallel work, 4 (shorter serial chain) otherwise
    uint32_t want_l0 = p.G >= (1ull << 18) ? 3u : 2u;
    p.l0 = c - 1 < want_l0 ? c - 1 : want_l0;
    p.m1 = p.B >> p.l0;
    p.nb0 = (p.m1 + (1u << H2_R0_LOG) - 1) >> H2_R0_LOG;
    uint32_t lm = ilog2_u32(p.m1);
    p.bits0 = lm < H2_R0_LOG ? lm : H2_R0_LOG;
    p.bits1 = ilog2_u32(p.nb0);
    p.r1_rows = 2 + p.bits0 + p.bits1;
}

// Batched-affine rounds (Msm::ba_round_body) for a planned one-shot MSM: `rounds` halvings of every work item's list of
// points, `target` pairs per thread (the batch one inversion is shared by).  Needs the binned layout with 8-aligned bins
// and work items (level r lives at index >> r), so forced odd capacities / item sizes (tests) switch it off.
#define H2_BA_MAX_ROUNDS 3
inline void msm_plan_ba(MsmPlan &p, uint32_t rounds, uint32_t target) {
    p.ba = 0;
    if (p.fixed || p.cap == 0 || (p.cap & 7u) || (p.cap_top & 7u) || (p.T & 7u) || rounds == 0) return;
    if (rounds > H2_BA_MAX_ROUNDS) rounds = H2_BA_MAX_ROUNDS;
    uint64_t mean = p.max_refs / (p.G ? p.G : 1);       // references per bucket
    if (mean > p.T) mean = p.T;
    if (mean < 4) return;                                // nothing to pair up
    p.ba = rounds;
    for (uint32_t r = 1; r <= rounds; r++) {
        // `target` pairs per thread in round 1; later rounds keep the thread count (half the pairs per inversion each time)
        // unless bit 16 of target asks for `target` pairs in every round (fewer threads in the later rounds)
        uint64_t pairs = mean >> ((target >> 16) & 1u ? r : 1u);
        if (pairs == 0) pairs = 1;
        uint64_t m = ((target & 0xffffu) + pairs / 2) / pairs;
        p.ba_m[r - 1] = (uint32_t)(m < 1 ? 1 : m > 64 ? 64 : m);
    }
}

struct MsmBuffers {
    // inputs
    const fe *scalars;        // n, canonical or Montgomery (see scalars_mont)
    const affine *bases;      // n, Montgomery coordinates
    affine *bases_phi;        // n, phi(bases) = (zeta x, y)   (GLV only)
    uint32_t scalars_mont;
    // scratch
    fe *scal_canon;           // n (only when scalars_mont)
    uint32_t *glv_parts;      // n x 8 words: |k1| (4 limbs, sign in bit 127) then |k2|   (GLV only)
    uint32_t *counts;         // G + 1  (histogram, then exclusive offsets after the scan)
    uint32_t *cursor;         // G   single-pass sort: bin fill (= bucket size)
    uint32_t *cursor2;        // G   exact sort: scatter cursor
    uint32_t *refs;           // max_refs   point index | sign << 31, sorted by bucket id
    uint32_t *size_hist;      // T + 2: [s] = number of work items of s references; [T + 1] = item count
    uint32_t *size_cursor;    // T + 1
    uint32_t *flags;          // [0] = some bucket exceeded T references (partials exist)
    uint2 *items;             // max_items  (bucket id, first reference), sorted by size descending
    xyzz *bucket_sum;         // G
    affine *ba[H2_BA_MAX_ROUNDS];   // batched-affine levels: ba[r - 1] holds (ref_space >> r) + 1 points, item lists at (start >> r)
    uint32_t *pkey, *pstart, *pend;   // part_total
    xyzz *ppt;                // part_total
    xyzz *ra_t, *ra_e;        // W * m1            level A: chunk totals / weighted sums
    xyzz *r0;                 // Wb * nb0 * 5      R0: per 8-entry block T, E, D_0..2
    xyzz *r1;                 // W * r1_rows       R1: per window T, E, D_0..
    xyzz *wsum;               // W                 2^(c w) S_w
    jacobian *result;         // 1 (sets in the batched fixed-base mode)
};

// Addition policies of the reduce kernels: a serial chain of XYZZ additions is latency-bound (5.5 us per add for a
// lone warp), so the device runs every chain on a QUAD of lanes (xyzz_add_quad: 4 multiply latencies instead of 14, same
// lane-multiplies); all four lanes hold the same values and execute the same loads / stores.
struct SerialAdd {
    template <class P> static H2_HD void add(xyzz &a, const xyzz &b) { xyzz_add<P>(a, b); }
    template <class P> static H2_HD void add_mixed(xyzz &a, const affine &b) { xyzz_add_mixed<P>(a, b); }
};
#ifdef __CUDACC__
struct QuadAdd {
    template <class P> static H2_D void add(xyzz &a, const xyzz &b) { xyzz_add_quad<P>(a, b); }
    template <class P> static H2_D void add_mixed(xyzz &a, const affine &b) { xyzz_add_mixed_quad<P>(a, b); }
};
struct PairAdd {
    template <class P> static H2_D void add_mixed(xyzz &a, const affine &b) { xyzz_add_mixed_pair<P>(a, b); }
};
#endif

// ---------------------------------------------------------------------------------------------
template <class P, class PS> struct Msm {
    // signed window digit; `carry` threads through the windows in increasing order
    static H2_HD int32_t next_digit(const uint32_t (&s)[8], uint32_t w, uint32_t c, uint32_t &carry) {
        uint32_t bit = w * c, idx = bit >> 5, sh = bit & 31;
        uint64_t v = 0;
        for (int i = 0; i < 8; i++) {   // register-resident select instead of dynamic indexing
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

    // GLV halves of scalar i: computed (and stored) by the first pass, re-loaded by the second
    static H2_HD void load_parts(const MsmPlan &p, const MsmBuffers &M, uint64_t i, bool first_pass, uint32_t (&part)[2][8], uint32_t (&neg)[2]) {
        neg[0] = neg[1] = 0;
        if (!p.glv) { uint32_t s[8]; load_scalar(M, i, s, first_pass); for (int k = 0; k < 8; k++) part[0][k] = s[k]; return; }
        uint4 *q = reinterpret_cast<uint4 *>(M.glv_parts) + i * 2;
        if (first_pass) {
            uint32_t s[8];
            load_scalar(M, i, s, true);
            glv_decompose<P>(s, part[0], neg[0], part[1], neg[1]);
            for (int e = 0; e < 2; e++)   // |k_e| < 2^127: four limbs, the sign rides in bit 127
                q[e] = make_uint4(part[e][0], part[e][1], part[e][2], part[e][3] | (neg[e] << 31));
        } else {
            for (int e = 0; e < 2; e++) {
                uint4 v = q[e];
                part[e][0] = v.x; part[e][1] = v.y; part[e][2] = v.z; part[e][3] = v.w & 0x7fffffffu; neg[e] = v.w >> 31;
                part[e][4] = part[e][5] = part[e][6] = part[e][7] = 0;
            }
        }
    }

    // Enumerates the (bucket, reference) pairs of scalar i: f(g, ref) for every non-zero digit.
    // ref = point index | endo << 30 (phi(P) instead of P) | negate << 31; fixed-base: index into the table.
    template <class F> static H2_HD bool for_each_digit(const MsmPlan &p, const MsmBuffers &M, uint64_t i, bool first_pass, F &&f) {
        uint32_t part[2][8], neg[2];
        const uint32_t halves = p.glv ? 2u : 1u;
        load_parts(p, M, i, first_pass, part, neg);
        bool ok = true;
        for (uint32_t e = 0; e < halves; e++) {
            uint32_t carry = 0;
            for (uint32_t w = 0; w < p.W; w++) {
                int32_t d = next_digit(part[e], w, p.c, carry);
                if (d == 0) continue;
                uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
                uint32_t g = (uint32_t)((uint64_t)(p.fixed ? i / p.n : w) * p.B + (mag - 1));
                uint32_t ref = (uint32_t)(p.fixed ? i % p.n + (uint64_t)w * p.stride : i);
                ref |= (e << 30) | ((((uint32_t)(d < 0)) ^ neg[e]) << 31);
                f(g, ref);
            }
            ok = ok && carry == 0;      // the top window must absorb the carry
        }
        return ok;
    }

    // ---- K4: work items.  Bucket g with cnt references yields ceil(cnt / T) items; item k covers
    // references [counts[g] + k T, min(+T, counts[g+1])).
    // Partial slot of a piece that does not cover its whole bucket: 2 * (start / T) + (k == 0).
    // Collision-free (a split bucket holds > T references, so two first-pieces never share a
    // T-aligned block, nor do two later pieces) and increasing with `start` (inside one block a
    // later piece of bucket g precedes the first piece of a bucket g' > g), which is what the merge
    // levels need: same-bucket partials in position order, no other key in between.
    static H2_HD uint64_t item_slot(const MsmPlan &p, uint32_t start, bool first_piece) {
        return 2ull * (start / p.T) + (first_piece ? 1 : 0);
    }
    // references of bucket g: refs[bucket_lo, bucket_hi).  flags[1] == 0: binned layout of the single-pass sort
    // (cursor[g] = size); flags[1] != 0: compact layout of the exact sort (counts = exclusive offsets).
    // bin of bucket g in the single-pass layout; false if g has no bin (a top-window digit beyond top_bins)
    static H2_HD bool bin_of(const MsmPlan &p, uint64_t g, uint32_t &lo, uint32_t &cap) {
        if (g < p.g_top) { lo = (uint32_t)(g * p.cap); cap = p.cap; return true; }
        const uint64_t b = g - p.g_top;
        lo = (uint32_t)(p.g_top * p.cap + b * p.cap_top); cap = p.cap_top;
        return b < p.top_bins;
    }
    static H2_HD uint32_t bucket_lo(const MsmPlan &p, const MsmBuffers &M, uint64_t g) {
        if (M.flags[1]) return M.counts[g];
        uint32_t lo, cap;
        return bin_of(p, g, lo, cap) ? lo : 0u;
    }
    static H2_HD uint32_t bucket_hi(const MsmPlan &p, const MsmBuffers &M, uint64_t g) {
        if (M.flags[1]) return M.counts[g + 1];
        uint32_t lo, cap;
        return bin_of(p, g, lo, cap) ? lo + M.cursor[g] : 0u;   // (a bucket without a bin is empty unless the sort overflowed)
    }
    // size histogram: thread per bucket
    static H2_HD void count_items(const MsmPlan &p, const MsmBuffers &M, uint64_t g, uint32_t &nfull, uint32_t &rem) {
        uint32_t cnt = bucket_hi(p, M, g) - bucket_lo(p, M, g);
        nfull = cnt / p.T; rem = cnt % p.T;
        if (cnt > p.T) M.flags[0] = 1;
    }
    // descending-size base offsets from the histogram (single thread): base[s] = #items larger than s
    static H2_HD void size_bases_body(const MsmPlan &p, const MsmBuffers &M) {
        uint32_t run = 0;
        for (uint32_t sz = p.T; sz >= 1; sz--) {
            uint32_t c = M.size_hist[sz];
            M.size_cursor[sz] = run;
            run += c;
        }
        M.size_cursor[0] = run;
        M.size_hist[p.T + 1] = run;     // total number of items
    }

    struct Flusher {
        const MsmBuffers *M;
        const MsmPlan *p;
        H2_HD void flush(uint32_t g, uint32_t a, uint32_t b, const xyzz &acc, uint64_t slot) {
            uint32_t lo = bucket_lo(*p, *M, g), hi = bucket_hi(*p, *M, g);
            if (a == lo && b == hi) {
                st_xyzz(M->bucket_sum + g, acc);
            } else {
                M->pkey[slot] = g; M->pstart[slot] = a; M->pend[slot] = b;
                st_xyzz(M->ppt + slot, acc);
            }
        }
    };

    // ---- K4a: batched-affine rounds.  An affine addition costs 3 multiplies (lambda = dy / dx, x3 = lambda^2 - x1 - x2,
    // y3 = lambda (x1 - x3) - y1) plus ONE inversion -- which Montgomery's trick shares among a whole batch of independent
    // additions at 3 more multiplies each: 6 per addition instead of the 10 of a mixed XYZZ addition.  The additions of a
    // bucket's chain are not independent, but the pairs of a halving round are: round r turns every work item's list of
    // len points into ceil(len / 2) sums (P0 + P1, P2 + P3, ..., an odd last point moves up unchanged).  Level 0 is the
    // item's references (gathered bases, negated / phi-mapped as the reference says), level r is stored at index >> r of
    // M.ba[r - 1] (bins and items are 8-aligned).  A thread takes ba_m[r - 1] consecutive items of the size-sorted item
    // list -- ~equal work per lane -- and runs its pairs in sub-batches of SUB: forward pass (denominators, running
    // product), one inversion by division steps (fe_inv_gcd, ~45 multiply-equivalents), backward pass (the sums).
    // Degenerate pairs keep the batch alive with a substitute denominator: an identity operand or P + (-P) uses 1, P + P
    // uses 2 y (the tangent slope 3 x^2 / 2 y).  After p.ba rounds accum0_pts_body runs the short XYZZ chain that is left.
    struct BaItem { uint32_t in_base, out_base, len, np; };      // len points at level r - 1, np = len / 2 pairs
    static H2_HD BaItem ba_item(const MsmPlan &p, const MsmBuffers &M, uint32_t r, uint64_t t) {
        const uint2 it = M.items[t];
        const uint32_t hi = bucket_hi(p, M, it.x), end = it.y + p.T < hi ? it.y + p.T : hi;
        BaItem b;
        b.len = (end - it.y + (1u << (r - 1)) - 1u) >> (r - 1);
        b.np = b.len >> 1; b.in_base = it.y >> (r - 1); b.out_base = it.y >> r;
        return b;
    }
    // Operand `key` of round r: a reference (r == 1: point index | phi << 30 | negate << 31) or a slot of level r - 1.
    static H2_HD const affine *ba_addr(const MsmPlan &p, const MsmBuffers &M, uint32_t r, uint32_t key) {
        if (r > 1) return M.ba[r - 2] + key;
        if (p.glv) return ((key >> 30) & 1u ? M.bases_phi : M.bases) + (key & 0x3fffffffu);
        return M.bases + (key & 0x7fffffffu);
    }
    static H2_HD affine ba_load(const MsmPlan &p, const MsmBuffers &M, uint32_t r, uint32_t key) {
        affine b = ld_affine(ba_addr(p, M, r, key));
        if (r == 1 && (key >> 31)) b.y = fe_neg<P>(b.y);
        return b;
    }
    static H2_HD uint32_t ba_key(const MsmBuffers &M, uint32_t r, uint32_t idx) { return r == 1 ? M.refs[idx] : idx; }
    static H2_HD void ba_prefetch(const void *q) {
#ifdef __CUDA_ARCH__
        asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
#else
        (void)q;
#endif
    }
    // kind of the addition A + B and its denominator: 0 generic (x2 - x1), 1 an operand is the identity (1),
    // 2 doubling (2 y1; y != 0 on a prime-order curve), 3 opposite points (1)
    static H2_HD uint32_t ba_kind(const affine &A, const affine &B, fe &d) {
        d = fe_sub<P>(B.x, A.x);
        if (!fe_is_zero(d) && !fe_is_zero(A.x) && !fe_is_zero(B.x)) return 0;
        if (affine_is_identity(A) || affine_is_identity(B)) { d = fe_one<P>(); return 1; }
        if (!fe_is_zero(d)) return 0;
        if (fe_eq(A.y, B.y)) { d = fe_dbl<P>(A.y); return 2; }
        d = fe_one<P>();
        return 3;
    }
    // One batch = the pairs [k_lo, k_hi) of item ta (single == true) or all pairs of items [ta, tb): forward pass
    // (denominators, running products pre[]), one inversion, backward pass (the sums).  Both passes walk the items in
    // chunks of U pairs: the chunk's 2 U operand keys are loaded first and their points prefetched, so 2 U gathers are in
    // flight per lane instead of one dependent reference -> point chain per pair (the kernel was bound by exactly that
    // latency: ncu r2i, long-scoreboard stalls at every use of a gathered value, 14 warps per SM).
    template <int U> static H2_HD void ba_batch(const MsmPlan &p, const MsmBuffers &M, uint32_t r, uint64_t ta, uint64_t tb, bool single,
                                                 uint32_t k_lo, uint32_t k_hi, fe *pre) {
        affine *out = M.ba[r - 1];
        fe acc = fe_one<P>();
        uint32_t c = 0;
        for (uint64_t tt = ta; tt < tb; tt++) {
            const BaItem it = ba_item(p, M, r, tt);
            const uint32_t ka = single ? k_lo : 0u, kb = single ? k_hi : it.np;
            for (uint32_t k = ka; k < kb; k += U) {
                uint32_t key[2 * U];
#pragma unroll
                for (int u = 0; u < 2 * U; u++) key[u] = 2 * k + u < 2 * kb ? ba_key(M, r, it.in_base + 2 * k + u) : 0u;
#pragma unroll
                for (int u = 0; u < 2 * U; u++) if (2 * k + u < 2 * kb) ba_prefetch(ba_addr(p, M, r, key[u]));
#pragma unroll
                for (int u = 0; u < U; u++) {
                    if (k + u >= kb) break;
                    const fe x1 = fe_load(&ba_addr(p, M, r, key[2 * u])->x), x2 = fe_load(&ba_addr(p, M, r, key[2 * u + 1])->x);
                    fe d = fe_sub<P>(x2, x1);
                    // the x coordinates alone decide the generic case; anything else looks at the whole points
                    if (fe_is_zero(d) || fe_is_zero(x1) || fe_is_zero(x2)) { affine A = ba_load(p, M, r, key[2 * u]), B = ba_load(p, M, r, key[2 * u + 1]); ba_kind(A, B, d); }
                    pre[c++] = acc;
                    acc = fe_mul_call<P>(acc, d);
                }
            }
        }
        if (c == 0) return;
        fe inv = fe_inv_gcd<P>(acc);
        // backward: inv = 1 / (d_0 ... d_c) on entry of step c
        for (uint64_t tt = tb; tt-- > ta;) {
            const BaItem it = ba_item(p, M, r, tt);
            const uint32_t ka = single ? k_lo : 0u, kb = single ? k_hi : it.np;
            if (kb <= ka) continue;
            for (uint32_t k = ka + ((kb - ka - 1) / U) * U;; k -= U) {
                uint32_t key[2 * U];
#pragma unroll
                for (int u = 0; u < 2 * U; u++) key[u] = 2 * k + u < 2 * kb ? ba_key(M, r, it.in_base + 2 * k + u) : 0u;
#pragma unroll
                for (int u = 0; u < 2 * U; u++) if (2 * k + u < 2 * kb) ba_prefetch(ba_addr(p, M, r, key[u]));
#pragma unroll
                for (int u = U - 1; u >= 0; u--) {
                    if (k + u >= kb) continue;
                    const affine A = ba_load(p, M, r, key[2 * u]), B = ba_load(p, M, r, key[2 * u + 1]);
                    fe d;
                    const uint32_t kind = ba_kind(A, B, d);
                    const fe dinv = fe_mul_call<P>(inv, pre[--c]);
                    inv = fe_mul_call<P>(inv, d);
                    affine S;
                    if (kind == 0 || kind == 2) {
                        fe num;
                        if (kind == 0) num = fe_sub<P>(B.y, A.y);
                        else { fe xx = fe_sqr_call<P>(A.x); num = fe_add<P>(fe_dbl<P>(xx), xx); }
                        const fe lam = fe_mul_call<P>(num, dinv);
                        S.x = fe_sub<P>(fe_sub<P>(fe_sqr_call<P>(lam), A.x), B.x);
                        S.y = fe_sub<P>(fe_mul_call<P>(lam, fe_sub<P>(A.x, S.x)), A.y);
                    } else if (kind == 1) {
                        S = affine_is_identity(A) ? B : A;
                    } else {
                        S.x = fe_zero(); S.y = fe_zero();
                    }
                    st_affine(out + it.out_base + k + u, S);
                }
                if (k == ka) break;
            }
        }
    }
    // PRE: pairs per inversion at most (the local array of running products); U: pairs per gather chunk (PRE % U == 0)
    template <int PRE = 128, int U = 4> static H2_HD void ba_round_body(const MsmPlan &p, const MsmBuffers &M, uint32_t r, uint64_t j) {
        if (M.flags[1]) return;                      // exact-sort layout: the classic kernel accumulates
        const uint64_t nitems = M.size_hist[p.T + 1];
        const uint32_t m = p.ba_m[r - 1];
        uint64_t t = j * m;
        const uint64_t t1 = t + m < nitems ? t + m : nitems;
        if (t >= nitems) return;
        affine *out = M.ba[r - 1];
        fe pre[PRE];
        while (t < t1) {
            // the next batch: consecutive items while their pairs fit PRE; an item with more pairs than that goes alone, in slices
            uint64_t tg = t;
            uint32_t total = 0;
            bool big = false;
            while (tg < t1) {
                const BaItem b = ba_item(p, M, r, tg);
                if (b.len & 1u) st_affine(out + b.out_base + b.np, ba_load(p, M, r, ba_key(M, r, b.in_base + 2 * b.np)));   // odd last point moves up
                if (b.np > (uint32_t)PRE) { big = tg == t; if (big) { total = b.np; tg++; } break; }
                if (total + b.np > (uint32_t)PRE) break;
                total += b.np; tg++;
            }
            if (big) for (uint32_t k0 = 0; k0 < total; k0 += PRE) ba_batch<U>(p, M, r, t, t + 1, true, k0, k0 + PRE < total ? k0 + PRE : total, pre);
            else ba_batch<U>(p, M, r, t, tg, false, 0, 0, pre);
            t = tg;
        }
    }
    // ... and the XYZZ chain over what the rounds left of item t (level p.ba), flushed like accum0_body does
    static H2_HD void accum0_pts_body(const MsmPlan &p, const MsmBuffers &M, uint64_t t) {
        if (M.flags[1] || t >= M.size_hist[p.T + 1]) return;
        const uint2 it = M.items[t];
        const uint32_t g = it.x, start = it.y, lo = bucket_lo(p, M, g), hi = bucket_hi(p, M, g);
        const uint32_t end = start + p.T < hi ? start + p.T : hi;
        const uint32_t len = (end - start + (1u << p.ba) - 1u) >> p.ba;
        const affine *src = M.ba[p.ba - 1] + (start >> p.ba);
        xyzz acc = xyzz_identity();
        for (uint32_t i = 0; i < len; i++) xyzz_add_mixed<P>(acc, ld_affine(src + i));
        Flusher F; F.M = &M; F.p = &p;
        F.flush(g, start, end, acc, p.part_offset[1] + item_slot(p, start, start == lo));
    }

    // level 0: one work item
    static H2_HD affine ref_point(const MsmPlan &p, const MsmBuffers &M, uint32_t ref) {
        if (p.glv) return ld_affine(((ref >> 30) & 1u ? M.bases_phi : M.bases) + (ref & 0x3fffffffu));
        return ld_affine(M.bases + (ref & 0x7fffffffu));
    }
    // AHEAD: the small-problem kernels (lanes cooperating on one item; the kernel lasts as long as its longest chain) fetch
    // the next reference and its point BEFORE the current addition, so the dependent reference -> point gather (~1.2 us) runs
    // under the addition's multiplies instead of between two of them.  The throughput kernel (one thread per item, 20 warps per
    // SM to hide latency, 96 registers) keeps the plain loop: 16 more live registers would cost it a CTA per SM.
    template <class MADD = SerialAdd, bool AHEAD = false> static H2_HD void accum0_body(const MsmPlan &p, const MsmBuffers &M, uint64_t t) {
        if (p.ba && !M.flags[1]) return;             // the batched-affine rounds + accum0_pts_body did the work
        if (p.fast && M.flags[1]) return;            // a bin overflowed and no exact sort follows: the references are not usable (the host re-runs)
        uint2 it;
        if (p.natural) {                             // item t = bucket t, whole (T >= the bin capacity in a fast pass)
            if (t >= p.G) return;
            it = make_uint2((uint32_t)t, bucket_lo(p, M, t));
            if (bucket_hi(p, M, t) == it.y) return;  // empty bucket: its sum stays the identity the memset wrote
        } else {
            if (t >= M.size_hist[p.T + 1]) return;
            it = M.items[t];
        }
        const uint32_t g = it.x, start = it.y, lo = bucket_lo(p, M, g), hi = bucket_hi(p, M, g);
        const uint32_t end = start + p.T < hi ? start + p.T : hi;
        xyzz acc = xyzz_identity();
        if (AHEAD) {
            uint32_t ref = M.refs[start];
            affine nxt = ref_point(p, M, ref);
            for (uint32_t pos = start; pos < end; pos++) {
                affine b = nxt;
                const uint32_t neg = ref >> 31;
                if (pos + 1 < end) { ref = M.refs[pos + 1]; nxt = ref_point(p, M, ref); }
                if (neg) b.y = fe_neg<P>(b.y);
                MADD::template add_mixed<P>(acc, b);
            }
        } else {
            for (uint32_t pos = start; pos < end; pos++) {
                const uint32_t ref = M.refs[pos];
                affine b = ref_point(p, M, ref);
                if (ref >> 31) b.y = fe_neg<P>(b.y);
                MADD::template add_mixed<P>(acc, b);
            }
        }
        Flusher F; F.M = &M; F.p = &p;
        F.flush(g, start, end, acc, p.part_offset[1] + item_slot(p, start, start == lo));
    }

    // level >= 1: chunk of partial slots; merged pieces go to 2 output slots per thread: slot 0 if
    // the piece lacks its bucket's start (it continues the previous thread's), slot 1 otherwise.
    static H2_HD void accumN_body(const MsmPlan &p, const MsmBuffers &M, uint32_t lv, uint64_t t) {
        if (!M.flags[0]) return;     // no bucket was split: nothing to merge
        const uint32_t K = p.acc_chunk[lv];
        const uint64_t in_base = p.part_offset[lv], slots = p.acc_slots[lv];
        uint64_t start = t * K;
        uint64_t end = start + K < slots ? start + K : slots;
        if (start >= end) return;
        Flusher F; F.M = &M; F.p = &p;
        const uint64_t out_base = (lv + 1 < p.acc_levels ? p.part_offset[lv + 1] : 0) + 2 * t;
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
                if (have) F.flush(cur, a, b, acc, out_base + (a > bucket_lo(p, M, cur) ? 0 : 1));
                have = true; cur = g; a = M.pstart[s]; b = M.pend[s]; acc = ld_xyzz(M.ppt + s);
            }
        }
        if (have) F.flush(cur, a, b, acc, out_base + (a > bucket_lo(p, M, cur) ? 0 : 1));
    }

    // ---- K5 level A: thread (w, u) reduces buckets [u L, u L + L) of window w:
    //   T = sum B[i],  E = sum (i - uL) B[i]    (running sums, no scalar multiplication)
    // ADD: SerialAdd (one thread per item; the host emulation) or QuadAdd (four lanes per item, device kernels)
    template <class ADD = SerialAdd> static H2_HD void reduceA_body(const MsmPlan &p, const MsmBuffers &M, uint64_t tid) {
        if (tid >= (uint64_t)p.Wb * p.m1) return;
        uint32_t w = (uint32_t)(tid / p.m1), u = (uint32_t)(tid % p.m1);
        const uint32_t L = 1u << p.l0;
        const xyzz *A = M.bucket_sum + (uint64_t)w * p.B + (uint64_t)u * L;
        xyzz run = xyzz_identity(), acc = xyzz_identity();
        if (p.chunks == 1 && p.G <= (1ull << 17)) {
            // small problems (latency-bound: a handful of warps on idle SMs): bucket i - 1 is fetched before the two additions of
            // bucket i, off the dependent chain (the throughput-bound large reduce keeps the plain loop: measured +1 % slower with it)
            xyzz nxt = ld_xyzz(A + (L - 1));
            for (uint32_t i = L - 1; i > 0; i--) {
                const xyzz cur = nxt;
                nxt = ld_xyzz(A + (i - 1));
                ADD::template add<P>(run, cur);
                ADD::template add<P>(acc, run);
            }
            ADD::template add<P>(run, nxt);
            uint64_t o1 = (uint64_t)w * p.m1 + u;
            st_xyzz(M.ra_t + o1, run);
            st_xyzz(M.ra_e + o1, acc);
            return;
        }
        for (uint32_t i = L - 1; i > 0; i--) {
            for (uint32_t k = 0; k < p.chunks; k++) ADD::template add<P>(run, ld_xyzz(A + k * p.G + i));
            ADD::template add<P>(acc, run);
        }
        for (uint32_t k = 0; k < p.chunks; k++) ADD::template add<P>(run, ld_xyzz(A + k * p.G));
        uint64_t o = (uint64_t)w * p.m1 + u;
        st_xyzz(M.ra_t + o, run);
        st_xyzz(M.ra_e + o, acc);
    }

    // Row semantics shared by R0 and R1: the value entry `idx` contributes to output row j.
    //   R0 (entries = level-A chunks of one 8-block): row 0 = T, row 1 = E, row 2+k = T if bit k of lane
    static H2_HD xyzz r0_contrib(const MsmPlan &p, const MsmBuffers &M, uint32_t w, uint32_t blk, uint32_t row, uint32_t lane) {
        uint32_t u = (blk << H2_R0_LOG) + lane;
        if (u >= p.m1) return xyzz_identity();
        uint64_t o = (uint64_t)w * p.m1 + u;
        if (row == 1) return ld_xyzz(M.ra_e + o);
        if (row >= 2 && !((lane >> (row - 2)) & 1u)) return xyzz_identity();
        return ld_xyzz(M.ra_t + o);
    }
    //   R1 (entries = R0 blocks of one window): rows 0..1+bits0 = plain sums of the R0 rows,
    //   row 2+bits0+k = R0 row 0 (T) of blocks whose index has bit k set
    static H2_HD xyzz r1_contrib(const MsmPlan &p, const MsmBuffers &M, uint32_t w, uint32_t row, uint32_t blk) {
        if (blk >= p.nb0) return xyzz_identity();
        const xyzz *e = M.r0 + ((uint64_t)w * p.nb0 + blk) * H2_R0_ROWS;
        const uint32_t plain = 2 + p.bits0;
        if (row < plain) return ld_xyzz(e + row);
        if (!((blk >> (row - plain)) & 1u)) return xyzz_identity();
        return ld_xyzz(e);
    }
    // 2^(c w) S_w with S_w = E + T + 2^l0 * sum_k 2^k D_k is a flat sum over the R1 rows of window w:
    // row r contributes R1[w][r] doubled c*w (+ k + l0 for the D_k row r = 2 + k) times.
    static H2_HD xyzz wsum_item(const MsmPlan &p, const MsmBuffers &M, uint32_t w, uint32_t r) {
        if (r >= p.r1_rows) return xyzz_identity();
        xyzz v = ld_xyzz(M.r1 + (uint64_t)w * p.r1_rows + r);
        uint32_t shift = (p.fixed ? 0 : p.c * w) + (r >= 2 ? (r - 2) + p.l0 : 0);   // fixed: w is a set index, not a window
        xyzz_shift<P>(v, shift);
        return v;
    }
    // Window table of a resident base set: table[w * stride + i] = affine(2^(c w) * base[i]).
    // One thread per base: W - 1 shifts of c Jacobian doublings each, then ONE inversion shared by
    // the thread's W - 1 points (Montgomery's trick) to bring them back to affine.  W <= 64 (c >= 4).
    static H2_HD void table_body(const affine *bases, affine *table, uint64_t count, uint64_t stride, uint32_t c, uint32_t W, uint64_t i) {
        if (i >= count) return;
        affine g = ld_affine(bases + i);
        st_affine(table + i, g);
        if (affine_is_identity(g) || W > 64) {
            for (uint32_t w = 1; w < W; w++) st_affine(table + (uint64_t)w * stride + i, g);   // identity stays identity
            return;
        }
        fe zs[64], pre[64];               // Z_w and Z_1 ... Z_(w-1)
        fe X = g.x, Y = g.y, Z = fe_one<P>(), run = fe_one<P>();
        for (uint32_t w = 1; w < W; w++) {
            for (uint32_t d = 0; d < c; d++) {      // dbl-2009-l, a = 0
                fe A = fe_sqr<P>(X), B = fe_sqr<P>(Y), C = fe_sqr<P>(B);
                fe t = fe_add<P>(X, B);
                fe D = fe_dbl<P>(fe_sub<P>(fe_sub<P>(fe_sqr<P>(t), A), C));
                fe E = fe_add<P>(fe_dbl<P>(A), A);
                fe F = fe_sqr<P>(E);
                fe Z3 = fe_dbl<P>(fe_mul<P>(Y, Z));
                X = fe_sub<P>(fe_sub<P>(F, D), D);
                Y = fe_sub<P>(fe_mul<P>(E, fe_sub<P>(D, X)), fe_dbl<P>(fe_dbl<P>(fe_dbl<P>(C))));
                Z = Z3;
            }
            affine tmp; tmp.x = X; tmp.y = Y;   // Jacobian X_w, Y_w parked in the slot
            st_affine(table + (uint64_t)w * stride + i, tmp);
            zs[w] = Z; pre[w] = run;
            run = fe_mul<P>(run, Z);
        }
        fe inv = fe_inv_gcd<P>(run);      // Z != 0: a prime-order curve has no 2-torsion
        for (uint32_t w = W - 1; w >= 1; w--) {
            fe zi = fe_mul<P>(inv, pre[w]);
            inv = fe_mul<P>(inv, zs[w]);
            fe zi2 = fe_sqr<P>(zi);
            affine a = ld_affine(table + (uint64_t)w * stride + i);
            a.x = fe_mul<P>(a.x, zi2);
            a.y = fe_mul<P>(a.y, fe_mul<P>(zi2, zi));
            st_affine(table + (uint64_t)w * stride + i, a);
        }
    }
    static H2_HD void finish(const MsmBuffers &M, const xyzz &total, uint32_t out_canonical, uint32_t set = 0) {
        jacobian j = xyzz_to_jacobian<P>(total);
        if (out_canonical) { j.x = fe_from_mont<P>(j.x); j.y = fe_from_mont<P>(j.y); j.z = fe_from_mont<P>(j.z); }
        st_jacobian(M.result + set, j);
    }
};

#if defined(__CUDACC__)
// "fetch and add 1" on per-lane addresses with a fast path for the skewed case: if every active
// lane of the warp hits the same counter (all-equal scalars, 0/1 columns), one lane adds the
// population count.  Returns the lane's slot (old value + rank) when WANT is set.
template <bool WANT> __device__ __forceinline__ uint32_t warp_inc(uint32_t *ctr, bool active) {
    const unsigned full = 0xffffffffu;
    unsigned mask = __ballot_sync(full, active);
    if (mask == 0) return 0;
    unsigned lane = threadIdx.x & 31u;
    int first = __ffs(mask) - 1;
    unsigned long long a0 = __shfl_sync(full, (unsigned long long)ctr, first);
    bool same = !active || (unsigned long long)ctr == a0;
    if (__all_sync(full, same)) {
        uint32_t base = 0;
        if ((int)lane == first) base = atomicAdd(ctr, (uint32_t)__popc(mask));
        if (WANT) base = __shfl_sync(full, base, first);
        return base + __popc(mask & ((1u << lane) - 1u));
    }
    if (!active) return 0;
    if (WANT) return atomicAdd(ctr, 1u);
    atomicAdd(ctr, 1u);
    return 0;
}

// digit loops are written out (not through for_each_digit) because the warp-synchronous atomics need
// every lane to execute every (half, window) step
// Single-pass sort: every (point, window) reference goes straight to its bucket's bin.  A full bin raises flags[1];
// the exact-sort kernels below then redo the job (they return at once otherwise).
template <class P, class PS> __global__ void __launch_bounds__(256) msm_bin_kernel(const MsmPlan p, const MsmBuffers M) {
    if (M.flags[1]) return;                      // exact sort requested by the host
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool in = i < p.n * p.sets;
    const uint64_t set = p.fixed && in ? i / p.n : 0, idx = p.fixed && in ? i % p.n : i;
    uint32_t part[2][8] = {{0, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0}}, sneg[2] = {0, 0};
    const uint32_t halves = p.glv ? 2u : 1u;
    if (in) Msm<P, PS>::load_parts(p, M, i, true, part, sneg);
    bool full = false;
    for (uint32_t e = 0; e < halves; e++) {
        uint32_t carry = 0;
        for (uint32_t w0 = 0; w0 < p.W; w0 += 4) {
            uint32_t slot[4], gid[4], ref[4];
            bool act[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t w = w0 + k;
                int32_t d = w < p.W ? Msm<P, PS>::next_digit(part[e], w, p.c, carry) : 0;
                act[k] = in && d != 0;
                uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
                gid[k] = (uint32_t)((uint64_t)(p.fixed ? set : (w < p.W ? w : 0)) * p.B + (act[k] ? mag - 1 : 0));
                ref[k] = ((uint32_t)idx + (uint32_t)(p.fixed && w < p.W ? (uint64_t)w * p.stride : 0)) | (e << 30) |
                         ((((uint32_t)(d < 0)) ^ sneg[e]) << 31);
                slot[k] = warp_inc<true>(M.cursor + gid[k], act[k]);
            }
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (act[k]) {
                    uint32_t lo, cap;
                    if (Msm<P, PS>::bin_of(p, gid[k], lo, cap) && slot[k] < cap) M.refs[lo + slot[k]] = ref[k];
                    else full = true;
                }
        }
    }
    if (full) M.flags[1] = 1;
}
template <class P, class PS> __global__ void __launch_bounds__(256) msm_hist_kernel(const MsmPlan p, const MsmBuffers M) {
    if (!M.flags[1]) return;                     // the single-pass sort succeeded
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool in = i < p.n * p.sets;
    const uint64_t set = p.fixed && in ? i / p.n : 0;
    uint32_t part[2][8] = {{0, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0}}, sneg[2] = {0, 0};
    const uint32_t halves = p.glv ? 2u : 1u;
    if (in) Msm<P, PS>::load_parts(p, M, i, true, part, sneg);
    for (uint32_t e = 0; e < halves; e++) {
        uint32_t carry = 0;
        for (uint32_t w = 0; w < p.W; w++) {
            int32_t d = Msm<P, PS>::next_digit(part[e], w, p.c, carry);
            bool act = in && d != 0;
            uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
            uint64_t g = (uint64_t)(p.fixed ? set : w) * p.B + (act ? mag - 1 : 0);
            warp_inc<false>(M.counts + g, act);
        }
    }
}
template <class P, class PS> __global__ void __launch_bounds__(256) msm_scatter_kernel(const MsmPlan p, const MsmBuffers M) {
    if (!M.flags[1]) return;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool in = i < p.n * p.sets;
    const uint64_t set = p.fixed && in ? i / p.n : 0, idx = p.fixed && in ? i % p.n : i;
    uint32_t part[2][8] = {{0, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0}}, sneg[2] = {0, 0};
    const uint32_t halves = p.glv ? 2u : 1u;
    if (in) Msm<P, PS>::load_parts(p, M, i, false, part, sneg);
    for (uint32_t e = 0; e < halves; e++) {
        uint32_t carry = 0;
        for (uint32_t w0 = 0; w0 < p.W; w0 += 4) {
            // four windows in flight so the returning atomics overlap
            uint32_t slot[4], gid[4], ref[4];
            bool act[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t w = w0 + k;
                int32_t d = w < p.W ? Msm<P, PS>::next_digit(part[e], w, p.c, carry) : 0;
                act[k] = in && d != 0;
                uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
                gid[k] = (uint32_t)((uint64_t)(p.fixed ? set : (w < p.W ? w : 0)) * p.B + (act[k] ? mag - 1 : 0));
                ref[k] = ((uint32_t)idx + (uint32_t)(p.fixed && w < p.W ? (uint64_t)w * p.stride : 0)) | (e << 30) |
                         ((((uint32_t)(d < 0)) ^ sneg[e]) << 31);
                slot[k] = warp_inc<true>(M.cursor2 + gid[k], act[k]);
            }
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (act[k]) M.refs[M.counts[gid[k]] + slot[k]] = ref[k];
        }
    }
}
// phi(P) = (zeta x, y) for the GLV split (the identity (0, 0) maps to itself)
template <class P, class PS> __global__ void __launch_bounds__(256) msm_phi_kernel(const affine *bases, affine *out, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    affine b = ld_affine(bases + i);
    b.x = fe_mul<P>(b.x, glv_zeta<P>());
    st_affine(out + i, b);
}
// work-item construction: size histogram, bases, then placement (descending size)
template <class P, class PS> __global__ void __launch_bounds__(256) msm_item_hist_kernel(const MsmPlan p, const MsmBuffers M) {
    if (p.fast && M.flags[1]) return;            // fast pass after a bin overflow: no exact sort follows, the counts are not there (the host re-runs)
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nfull = 0, rem = 0;
    if (g < p.G) Msm<P, PS>::count_items(p, M, g, nfull, rem);
    if (nfull) atomicAdd(M.size_hist + p.T, nfull);
    warp_inc<false>(M.size_hist + rem, rem != 0);
}
template <class P, class PS> __global__ void msm_item_bases_kernel(const MsmPlan p, const MsmBuffers M) {
    if (threadIdx.x == 0 && blockIdx.x == 0) Msm<P, PS>::size_bases_body(p, M);
}
template <class P, class PS> __global__ void __launch_bounds__(256) msm_item_place_kernel(const MsmPlan p, const MsmBuffers M) {
    if (p.fast && M.flags[1]) return;
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nfull = 0, rem = 0, lo = 0;
    if (g < p.G) { Msm<P, PS>::count_items(p, M, g, nfull, rem); lo = Msm<P, PS>::bucket_lo(p, M, g); }
    if (nfull) {
        uint32_t at = atomicAdd(M.size_cursor + p.T, nfull);
        for (uint32_t k = 0; k < nfull; k++) M.items[at + k] = make_uint2((uint32_t)g, lo + k * p.T);
    }
    uint32_t at = warp_inc<true>(M.size_cursor + rem, rem != 0);
    if (rem) M.items[at] = make_uint2((uint32_t)g, lo + nfull * p.T);
}
template <class P, class PS> __global__ void __launch_bounds__(128) msm_table_kernel(const affine *bases, affine *table, uint64_t count, uint64_t stride,
                                                                                    uint32_t c, uint32_t W) {
    Msm<P, PS>::table_body(bases, table, count, stride, c, W, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
template <class P, class PS> __global__ void __launch_bounds__(128, 5) msm_accum0_kernel(const MsmPlan p, const MsmBuffers M) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    Msm<P, PS>::accum0_body(p, M, t);
}
// batched-affine round r (Msm::ba_round_body) and the XYZZ chain over its last level
template <class P, class PS, int U, int MINB> __global__ void __launch_bounds__(128, MINB) msm_ba_round_kernel(const MsmPlan p, const MsmBuffers M, uint32_t r) {
    Msm<P, PS>::template ba_round_body<128, U>(p, M, r, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
template <class P, class PS> __global__ void __launch_bounds__(128, 5) msm_accum0_pts_kernel(const MsmPlan p, const MsmBuffers M) {
    Msm<P, PS>::accum0_pts_body(p, M, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
// small problems: one QUAD per work item -- the kernel lasts as long as the longest bucket's chain of additions, and a
// quad runs that chain 2.5x faster (4 multiply latencies per mixed addition instead of 10)
template <class P, class PS> __global__ void __launch_bounds__(128) msm_accum0_quad_kernel(const MsmPlan p, const MsmBuffers M) {
    uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    Msm<P, PS>::template accum0_body<QuadAdd, true>(p, M, t);
}
// ... or one PAIR of lanes per work item (xyzz_add_mixed_pair: no idle multiply slots, 5 multiply latencies per addition)
template <class P, class PS> __global__ void __launch_bounds__(128) msm_accum0_pair_kernel(const MsmPlan p, const MsmBuffers M) {
    uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    Msm<P, PS>::template accum0_body<PairAdd, true>(p, M, t);
}
// ... and WAYS quads per work item (test hook h2_test_set_accum_ways; NOT the default: measured at k = 14, c = 15 -- ~17
// references per bucket, the fullest ~35 -- 2 / 4 ways change a commit by -9 % / 0 % and the IPA opening by +6 % / +32 %: the
// quad accumulation is bound by lane-multiplies, not by its chains): quad h of the group adds references start + h, start + h + WAYS, ...; the
// partial sums are then folded with log2(WAYS) shuffle + quad-add steps and quad 0 writes the result.  All lanes of a
// group share the work item, so the group leaves together; the quads of a group run different trip counts and meet again
// at the group-wide shuffles.
template <class P, class PS, int WAYS> __global__ void __launch_bounds__(128) msm_accum0_multi_kernel(const MsmPlan p, const MsmBuffers M) {
    const uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / (4 * WAYS);
    if (p.fast && M.flags[1]) return;
    if (t >= M.size_hist[p.T + 1]) return;
    const uint32_t lane = threadIdx.x & 31u, h = (lane >> 2) & (WAYS - 1);
    const uint32_t gmask = (WAYS == 8 ? 0xffffffffu : ((1u << (4 * WAYS)) - 1u) << (lane & ~(4u * WAYS - 1u)));
    const uint2 it = M.items[t];
    const uint32_t g = it.x, start = it.y, lo = Msm<P, PS>::bucket_lo(p, M, g), hi = Msm<P, PS>::bucket_hi(p, M, g);
    const uint32_t end = start + p.T < hi ? start + p.T : hi;
    xyzz acc = xyzz_identity();
    for (uint32_t pos = start + h; pos < end; pos += WAYS) {
        const uint32_t ref = M.refs[pos];
        affine b;
        if (p.glv) b = ld_affine(((ref >> 30) & 1u ? M.bases_phi : M.bases) + (ref & 0x3fffffffu));
        else b = ld_affine(M.bases + (ref & 0x7fffffffu));
        if (ref >> 31) b.y = fe_neg<P>(b.y);
        xyzz_add_mixed_quad<P>(acc, b);
    }
#pragma unroll
    for (int step = WAYS / 2; step >= 1; step >>= 1) {
        xyzz other;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            other.x.v[i] = __shfl_down_sync(gmask, acc.x.v[i], 4 * step, 4 * WAYS);
            other.y.v[i] = __shfl_down_sync(gmask, acc.y.v[i], 4 * step, 4 * WAYS);
            other.zz.v[i] = __shfl_down_sync(gmask, acc.zz.v[i], 4 * step, 4 * WAYS);
            other.zzz.v[i] = __shfl_down_sync(gmask, acc.zzz.v[i], 4 * step, 4 * WAYS);
        }
        if ((int)h < step) xyzz_add_quad<P>(acc, other);
    }
    if (h == 0) {
        typename Msm<P, PS>::Flusher F; F.M = &M; F.p = &p;
        F.flush(g, start, end, acc, p.part_offset[1] + Msm<P, PS>::item_slot(p, start, start == lo));
    }
}
// ... or WAYS independent LANES per work item: lane h adds references start + h, start + h + WAYS, ... with the plain serial
// mixed addition (10 multiplies, no exchange between lanes, next operand fetched ahead), then log2(WAYS) shuffle + full-addition
// steps fold the partial sums.  A cooperative addition is no shorter than a serial one in practice (a level of the pair / quad
// forms costs its multiply plus ~0.3-0.7 us of shuffles and selects: 3.3-4.4 us per addition against 3.6 us serial), so what
// shortens the longest bucket's chain is cutting it into WAYS independent pieces.
template <class P, class PS, int WAYS> __global__ void __launch_bounds__(128) msm_accum0_split_kernel(const MsmPlan p, const MsmBuffers M) {
    const uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / WAYS;
    if (p.fast && M.flags[1]) return;
    const uint32_t lane = threadIdx.x & 31u, h = lane & (WAYS - 1);
    uint2 it;
    bool live = true;
    if (p.natural) {
        live = t < p.G;
        it = make_uint2((uint32_t)(live ? t : 0), live ? Msm<P, PS>::bucket_lo(p, M, live ? t : 0) : 0u);
    } else {
        live = t < M.size_hist[p.T + 1];
        it = live ? M.items[t] : make_uint2(0u, 0u);
    }
    // lanes of a warp leave together (the shuffles below are warp-wide): a dead group just carries identities
    const uint32_t g = it.x, start = it.y, lo = Msm<P, PS>::bucket_lo(p, M, g), hi = live ? Msm<P, PS>::bucket_hi(p, M, g) : start;
    const uint32_t end = start + p.T < hi ? start + p.T : hi;
    xyzz acc = xyzz_identity();
    if (start + h < end) {
        uint32_t ref = M.refs[start + h];
        affine nxt = Msm<P, PS>::ref_point(p, M, ref);
        for (uint32_t pos = start + h; pos < end; pos += WAYS) {
            affine b = nxt;
            const uint32_t neg = ref >> 31;
            if (pos + WAYS < end) { ref = M.refs[pos + WAYS]; nxt = Msm<P, PS>::ref_point(p, M, ref); }
            if (neg) b.y = fe_neg<P>(b.y);
            xyzz_add_mixed<P>(acc, b);
        }
    }
#pragma unroll
    for (int step = WAYS / 2; step >= 1; step >>= 1) {
        xyzz other;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            other.x.v[i] = __shfl_down_sync(0xffffffffu, acc.x.v[i], step, WAYS);
            other.y.v[i] = __shfl_down_sync(0xffffffffu, acc.y.v[i], step, WAYS);
            other.zz.v[i] = __shfl_down_sync(0xffffffffu, acc.zz.v[i], step, WAYS);
            other.zzz.v[i] = __shfl_down_sync(0xffffffffu, acc.zzz.v[i], step, WAYS);
        }
        if ((int)h < step) xyzz_add<P>(acc, other);
    }
    if (h == 0 && live && end > start) {
        typename Msm<P, PS>::Flusher F; F.M = &M; F.p = &p;
        F.flush(g, start, end, acc, p.part_offset[1] + Msm<P, PS>::item_slot(p, start, start == lo));
    }
}
template <class P, class PS> __global__ void __launch_bounds__(128) msm_accumN_kernel(const MsmPlan p, const MsmBuffers M, uint32_t lv) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < p.acc_threads[lv]) Msm<P, PS>::accumN_body(p, M, lv, t);
}
// Partial-merge levels >= 3 in a single CTA (empty unless some bucket exceeded T references)
template <class P, class PS> __global__ void __launch_bounds__(256) msm_accum_rest_kernel(const MsmPlan p, const MsmBuffers M) {
    if (!M.flags[0]) return;
    for (uint32_t lv = 3; lv < p.acc_levels; lv++) {
        for (uint64_t t = threadIdx.x; t < p.acc_threads[lv]; t += blockDim.x) Msm<P, PS>::accumN_body(p, M, lv, t);
        __threadfence();
        __syncthreads();
    }
}
template <class P, class PS> __global__ void __launch_bounds__(128) msm_reduceA_kernel(const MsmPlan p, const MsmBuffers M) {
    uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;       // one quad per chunk
    Msm<P, PS>::template reduceA_body<QuadAdd>(p, M, t);
}

// Shared-memory tree sum of one value per thread over the `width` (power of two) lanes of a row;
// rows are laid out [row][lane].  Every thread of the CTA must call it.  Total lands in lane 0.
template <class P> __device__ __forceinline__ void block_tree_sum(xyzz *sh, xyzz &v, uint32_t row, uint32_t lane, uint32_t width) {
    st_xyzz(sh + row * width + lane, v);
    __syncthreads();
    for (uint32_t off = width >> 1; off > 0; off >>= 1) {
        if (lane < off) {
            xyzz o = ld_xyzz(sh + row * width + lane + off);
            xyzz_add<P>(v, o);
            st_xyzz(sh + row * width + lane, v);
        }
        __syncthreads();
    }
}
// The same for quads: entry `qd` is held by the four lanes of quad qd; every thread of the CTA must call it.
template <class P> __device__ __forceinline__ void quad_tree_sum(xyzz *sh, xyzz &v, uint32_t qd, uint32_t quads) {
    st_xyzz(sh + qd, v);
    __syncthreads();
    for (uint32_t off = quads >> 1; off > 0; off >>= 1) {
        if (qd < off) {
            xyzz o = ld_xyzz(sh + qd + off);
            xyzz_add_quad<P>(v, o);
            st_xyzz(sh + qd, v);
        }
        __syncthreads();
    }
}
// R0: one QUAD per (window, 8-block, row): a short serial sum keeps every quad busy (a shared-memory
// tree runs its adds at 1/2 .. 1/32 lane utilisation).
template <class P, class PS> __global__ void __launch_bounds__(128) msm_r0_kernel(const MsmPlan p, const MsmBuffers M) {
    uint64_t tid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const uint32_t rows = 2 + p.bits0;
    if (tid >= (uint64_t)p.Wb * p.nb0 * rows) return;
    uint32_t row = (uint32_t)(tid % rows);
    uint64_t wb = tid / rows;
    uint32_t w = (uint32_t)(wb / p.nb0), blk = (uint32_t)(wb % p.nb0);
    xyzz v = xyzz_identity();
    // the contributing lanes of this row, each operand fetched one addition ahead of its use
    const uint32_t want = row >= 2 ? 1u << (row - 2) : 0u;
    auto next_lane = [&](uint32_t l) { while (l < (1u << H2_R0_LOG) && (l & want) != want) l++; return l; };
    uint32_t lane = next_lane(0);
    xyzz nxt = lane < (1u << H2_R0_LOG) ? Msm<P, PS>::r0_contrib(p, M, w, blk, row, lane) : xyzz_identity();
    while (lane < (1u << H2_R0_LOG)) {
        const xyzz c = nxt;
        lane = next_lane(lane + 1);
        if (lane < (1u << H2_R0_LOG)) nxt = Msm<P, PS>::r0_contrib(p, M, w, blk, row, lane);
        xyzz_add_quad<P>(v, c);
    }
    st_xyzz(M.r0 + ((uint64_t)w * p.nb0 + blk) * H2_R0_ROWS + row, v);
}
// R1: one CTA per (window, output row): 128 quads stride over the window's nb0 blocks, then a tree over the quads
#define H2_R1_QUADS 128
template <class P, class PS> __global__ void __launch_bounds__(4 * H2_R1_QUADS) msm_r1_kernel(const MsmPlan p, const MsmBuffers M) {
    __shared__ xyzz sh[H2_R1_QUADS];
    const uint32_t w = blockIdx.x / p.r1_rows, row = blockIdx.x % p.r1_rows, qd = threadIdx.x >> 2;
    xyzz v = xyzz_identity();
    xyzz nxt = qd < p.nb0 ? Msm<P, PS>::r1_contrib(p, M, w, row, qd) : xyzz_identity();
    for (uint32_t blk = qd; blk < p.nb0; blk += H2_R1_QUADS) {
        const xyzz c = nxt;
        if (blk + H2_R1_QUADS < p.nb0) nxt = Msm<P, PS>::r1_contrib(p, M, w, row, blk + H2_R1_QUADS);
        xyzz_add_quad<P>(v, c);
    }
    quad_tree_sum<P>(sh, v, qd, H2_R1_QUADS);
    if (threadIdx.x == 0) st_xyzz(M.r1 + (uint64_t)w * p.r1_rows + row, v);
}
// Window value 2^(c w) S_w: one CTA per window, one QUAD of lanes per R1 row (<= 32 rows) for the shift
// (xyzz_shift_quad), then a shared-memory tree sum over the rows.  (Msm::wsum_item is the serial form the host
// emulation runs.)
template <class P, class PS> __global__ void __launch_bounds__(128) msm_wsum_kernel(const MsmPlan p, const MsmBuffers M) {
    __shared__ xyzz sh[32];
    const uint32_t w = blockIdx.x, r = threadIdx.x >> 2;
    xyzz v = xyzz_identity();
    if (r < p.r1_rows) {
        v = ld_xyzz(M.r1 + (uint64_t)w * p.r1_rows + r);
        uint32_t shift = (p.fixed ? 0 : p.c * w) + (r >= 2 ? (r - 2) + p.l0 : 0);   // fixed: w is a set index, not a window
        xyzz_shift_quad<P>(v, shift);
    }
    quad_tree_sum<P>(sh, v, r, 32);
    if (threadIdx.x == 0) st_xyzz(M.wsum + w, v);
}
// Final: tree sum of the W window values (16 quads)
template <class P, class PS> __global__ void __launch_bounds__(64) msm_final_kernel(const MsmPlan p, const MsmBuffers M, uint32_t out_canonical) {
    __shared__ xyzz sh[16];
    if (p.fixed) {   // one result per set (the sets are independent MSMs)
        for (uint32_t set = threadIdx.x; set < p.sets; set += 64) Msm<P, PS>::finish(M, ld_xyzz(M.wsum + set), out_canonical, set);
        return;
    }
    const uint32_t qd = threadIdx.x >> 2;
    xyzz v = xyzz_identity();
    for (uint32_t w = qd; w < p.Wb; w += 16) { xyzz c = ld_xyzz(M.wsum + w); xyzz_add_quad<P>(v, c); }
    quad_tree_sum<P>(sh, v, qd, 16);
    if (threadIdx.x == 0) Msm<P, PS>::finish(M, v, out_canonical);
}

// exclusive scan of counts[0..G] in place (counts[G] becomes the total), three small kernels
#define H2_SCAN_BLOCK 1024
#define H2_SCAN_ITEMS 8
static __global__ void __launch_bounds__(H2_SCAN_BLOCK) scan_block_sums_kernel(const uint32_t *in, uint64_t n, uint32_t *block_sums, const uint32_t *only_if) {
    __shared__ uint32_t sh[32];
    if (only_if && !*only_if) return;
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
static __global__ void __launch_bounds__(H2_SCAN_BLOCK) scan_single_block_kernel(uint32_t *a, uint32_t n, const uint32_t *only_if) {
    // exclusive scan of a[0..n) by one block, n arbitrary (loops in tiles of blockDim)
    __shared__ uint32_t sh[H2_SCAN_BLOCK];
    __shared__ uint32_t carry_s;
    if (only_if && !*only_if) return;
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
static __global__ void __launch_bounds__(H2_SCAN_BLOCK) scan_apply_kernel(uint32_t *a, uint64_t n, const uint32_t *block_offsets, const uint32_t *only_if) {
    __shared__ uint32_t sh[H2_SCAN_BLOCK];
    if (only_if && !*only_if) return;
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
