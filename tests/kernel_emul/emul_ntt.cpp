// TEST-ONLY serial execution of the NTT kernel phases (ntt.cuh) on the host.
#include <cstring>
#include <vector>
#include "ntt.cuh"
using namespace h2;

// mode 0: plain best_fft; 1: ifft (out_scale = divisor); 2: coeff_to_extended (in_log_n = k, in_scale zeta
// powers); 3: extended_to_coeff (out_scale = divisor * [1, z^2, z], truncate to out_len).
// All elements canonical bytes.  tile_log/max_sp override the plan (0 = defaults) to exercise more geometries.
template <class P>
static int run_ntt(int mode, const uint8_t *in, uint32_t in_log_n, uint32_t log_n, const uint8_t *omega,
                   const uint8_t *zeta, const uint8_t *divisor, uint64_t out_len, uint8_t *out, uint32_t nthr) {
    uint64_t n = 1ull << log_n, n_in = 1ull << in_log_n;
    std::vector<fe> a(n_in), work(n), res(n), tw(n / 2 ? n / 2 : 1), pow2(32);
    fe w, z = fe_zero(), dv = fe_zero();
    memcpy(w.v, omega, 32); w = fe_to_mont<P>(w);
    if (zeta) { memcpy(z.v, zeta, 32); z = fe_to_mont<P>(z); }
    if (divisor) { memcpy(dv.v, divisor, 32); dv = fe_to_mont<P>(dv); }
    for (uint64_t i = 0; i < n_in; i++) memcpy(a[i].v, in + 32 * i, 32);   // canonical in
    TwiddleGen<P>::pow2_body(pow2.data(), w, log_n ? log_n : 1);
    for (uint64_t t = 0; t * 32 < (n / 2); t++) TwiddleGen<P>::fill_body(tw.data(), pow2.data(), n / 2, t);
    uint32_t sp[8], logc[8];
    int passes = ntt_plan(log_n, sp, logc);
    fe R2 = fe_r2<P>(), one_c = fe_zero(); one_c.v[0] = 1;
    uint32_t s0 = 0;
    for (int i = 0; i < passes; i++) {
        NttPassArgs A;
        A.in = i == 0 ? a.data() : work.data();
        A.out = i == passes - 1 ? res.data() : work.data();
        A.tw = tw.data(); A.log_n = log_n; A.s0 = s0; A.sp = sp[i]; A.logc = logc[i];
        A.flags = (i == 0 ? NTT_FIRST | NTT_IN_SCALE : 0) | (i == passes - 1 ? NTT_LAST | NTT_OUT_SCALE : 0);
        A.in_log_n = in_log_n; A.out_len = out_len;
        // canonical -> Montgomery folded into in_scale: mont_mul(a_canon, c * R^2) = a c R
        fe zp[3] = {fe_one<P>(), z, fe_mul<P>(z, z)};
        for (int k = 0; k < 3; k++) A.in_scale[k] = (mode == 2) ? fe_mul<P>(zp[k], R2) : R2;
        // Montgomery -> canonical folded into out_scale: mont_mul(x R, c) = x c
        fe oc[3] = {one_c, one_c, one_c};
        if (mode == 1) for (int k = 0; k < 3; k++) oc[k] = fe_from_mont<P>(dv);
        if (mode == 3) { oc[0] = fe_from_mont<P>(dv); oc[1] = fe_from_mont<P>(fe_mul<P>(dv, zp[2])); oc[2] = fe_from_mont<P>(fe_mul<P>(dv, zp[1])); }
        for (int k = 0; k < 3; k++) A.out_scale[k] = oc[k];
        uint32_t tiles = (uint32_t)(n >> (sp[i] + logc[i]));
        std::vector<uint4> sm(ntt_smem_bytes(sp[i], logc[i]) / 16), twc(ntt_twc_bytes(sp[i], logc[i], i == passes - 1) / 16 + 1);
        const bool last = i == passes - 1;
        if ((nthr & 1) && NttDense<P>::supported(A)) {
            // odd thread counts select the dense-layout path of ntt_pass_tma_kernel: the bulk copies become memcpy
            std::vector<uint8_t> buf(ntt_tma_buf_bytes(sp[i], logc[i], last) + 64), outst((ntt_tma_rowb(logc[i]) << sp[i]) + 64);
            const fe *src = A.in;
            std::vector<fe> dst_tmp;
            fe *dst = A.out;
            if (A.in == A.out) { dst_tmp.assign(A.in, A.in + n); src = dst_tmp.data(); }   // the kernel's in-place passes read a tile before writing it
            for (uint32_t tile = 0; tile < tiles; tile++) {
                for (uint32_t u = 0; u < NttDense<P>::in_units(A); u++) {
                    auto spn = NttDense<P>::in_span(A, tile, u);
                    if (spn.valid) memcpy(buf.data() + spn.smem_off, src + spn.elem, spn.bytes); else memset(buf.data() + spn.smem_off, 0, spn.bytes);
                }
                for (uint32_t t = 0; t < nthr; t++) NttPass<P>::twiddle_phase(A, tile, t, nthr, twc.data());
                typename NttDense<P>::Layout lay;
                lay.buf = buf.data(); lay.out = last ? outst.data() : buf.data();
                lay.rowb = ntt_tma_rowb(logc[i]); lay.colb = ntt_tma_colb(sp[i]); lay.geomB = last;
                for (uint32_t st = 0; st < NttPass<P>::num_steps(sp[i]); st++)
                    for (uint32_t t = 0; t < nthr; t++) NttPass<P>::step_phase_l(A, tile, st, t, nthr, lay, twc.data());
                for (uint32_t r = 0; r < (1u << sp[i]); r++) {
                    auto spn = NttDense<P>::out_span(A, tile, r);
                    if (spn.valid) memcpy(dst + spn.elem, lay.out + spn.smem_off, spn.bytes);
                }
            }
        } else
        for (uint32_t tile = 0; tile < tiles; tile++) {      // the same phase order as ntt_pass_kernel, one barrier between phases
            for (uint32_t t = 0; t < nthr; t++) NttPass<P>::twiddle_phase(A, tile, t, nthr, twc.data());
            for (uint32_t t = 0; t < nthr; t++) NttPass<P>::load_phase(A, tile, t, nthr, sm.data());
            if (tile & 1) {                                   // the one-stage-at-a-time form stays as a cross-check of the radix-4 steps
                for (uint32_t sl = 1; sl <= sp[i]; sl++)
                    for (uint32_t t = 0; t < nthr; t++) NttPass<P>::stage_phase(A, tile, sl, t, nthr, sm.data());
            } else {
                for (uint32_t st = 0; st < NttPass<P>::num_steps(sp[i]); st++)
                    for (uint32_t t = 0; t < nthr; t++) NttPass<P>::step_phase(A, tile, st, t, nthr, sm.data(), twc.data());
            }
            for (uint32_t t = 0; t < nthr; t++) NttPass<P>::store_phase(A, tile, t, nthr, sm.data());
        }
        s0 += sp[i];
    }
    uint64_t outn = out_len < n ? out_len : n;
    for (uint64_t i = 0; i < outn; i++) memcpy(out + 32 * i, res[i].v, 32);
    return passes;
}
extern "C" int emu_ntt(int field, int mode, const uint8_t *in, uint32_t in_log_n, uint32_t log_n, const uint8_t *omega,
                       const uint8_t *zeta, const uint8_t *divisor, uint64_t out_len, uint8_t *out, uint32_t nthr) {
    if (field == 0) return run_ntt<FpParams>(mode, in, in_log_n, log_n, omega, zeta, divisor, out_len, out, nthr);
    return run_ntt<FqParams>(mode, in, in_log_n, log_n, omega, zeta, divisor, out_len, out, nthr);
}
