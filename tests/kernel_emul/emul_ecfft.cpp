// TEST-ONLY serial execution of the EC-FFT / batch-normalise kernel bodies (ecfft.cuh) on the host.
#include <cstring>
#include <vector>
#include "ecfft.cuh"
using namespace h2;

// mode 0: h2_ec_fft (Jacobian in -> Jacobian out, optional scale); mode 1: h2_params_lagrange (affine in -> EC-FFT ->
// scale -> batch_normalize -> affine out).  All elements canonical bytes.
template <class P, class PS>
static int run_ecfft(int mode, int quad, const uint8_t *in, uint32_t log_n, const uint8_t *omega, const uint8_t *scale, uint8_t *out) {
    const uint64_t n = 1ull << log_n;
    std::vector<xyzz> work(n);
    std::vector<fe> tw(n / 2 ? n / 2 : 1), pow2(64);
    fe w; memcpy(w.v, omega, 32); w = fe_to_mont<PS>(w);
    TwiddleGen<PS>::pow2_body(pow2.data(), w, log_n ? log_n : 1);
    for (uint64_t t = 0; t * 32 < (n / 2 ? n / 2 : 1); t++) TwiddleGen<PS>::fill_body(tw.data(), pow2.data(), n / 2 ? n / 2 : 1, t);
    if (mode == 0) {
        std::vector<jacobian> a(n);
        memcpy(a.data(), in, n * sizeof(jacobian));
        for (uint64_t j = 0; j < n; j++) EcFft<P, PS>::load_jac_body(a.data(), 1, work.data(), log_n, j);
    } else {
        std::vector<affine> a(n);
        memcpy(a.data(), in, n * sizeof(affine));
        for (uint64_t j = 0; j < n; j++) EcFft<P, PS>::load_affine_body(a.data(), 1, work.data(), log_n, j);
    }
    for (uint32_t s = 1; s <= log_n; s++)
        for (uint64_t t = 0; t < n / 2; t++) {
            if (quad) EcFft<P, PS>::stage_body_q(work.data(), tw.data(), log_n, s, t, true);
            else EcFft<P, PS>::stage_body(work.data(), tw.data(), log_n, s, t);
        }
    if (scale) {
        fe sc; memcpy(sc.v, scale, 32);
        for (uint64_t i = 0; i < n; i++) {
            if (quad) EcFft<P, PS>::scale_body_q(work.data(), sc, i, true);
            else EcFft<P, PS>::scale_body(work.data(), sc, i);
        }
    }
    if (mode == 0) {
        std::vector<jacobian> o(n);
        for (uint64_t i = 0; i < n; i++) EcFft<P, PS>::store_jac_body(work.data(), o.data(), 1, i);
        memcpy(out, o.data(), n * sizeof(jacobian));
    } else {
        std::vector<affine> o(n);
        for (uint64_t t = 0; t * H2_NORM_CHUNK < n; t++) Normalize<P>::body(work.data(), nullptr, 0, o.data(), 1, n, t);
        memcpy(out, o.data(), n * sizeof(affine));
    }
    return 0;
}
extern "C" int emu_ec_fft(int curve, int mode, int quad, const uint8_t *in, uint32_t log_n, const uint8_t *omega, const uint8_t *scale, uint8_t *out) {
    if (curve == 0) return run_ecfft<FpParams, FqParams>(mode, quad, in, log_n, omega, scale, out);
    return run_ecfft<FqParams, FpParams>(mode, quad, in, log_n, omega, scale, out);
}
template <class P> static int run_norm(const uint8_t *in_xyz, uint64_t n, uint8_t *out_xy) {
    std::vector<jacobian> a(n ? n : 1);
    std::vector<affine> o(n ? n : 1);
    memcpy(a.data(), in_xyz, n * sizeof(jacobian));
    for (uint64_t t = 0; t * H2_NORM_CHUNK < n; t++) Normalize<P>::body(nullptr, a.data(), 1, o.data(), 1, n, t);
    memcpy(out_xy, o.data(), n * sizeof(affine));
    return 0;
}
extern "C" int emu_batch_normalize(int curve, const uint8_t *in_xyz, uint64_t n, uint8_t *out_xy) {
    return curve == 0 ? run_norm<FpParams>(in_xyz, n, out_xy) : run_norm<FqParams>(in_xyz, n, out_xy);
}
// k * b through the GLV joint ladder; b affine canonical, k canonical 32 bytes; out affine canonical
template <class P> static int run_glv_mul(int quad, const uint8_t *b_xy, const uint8_t *k, uint8_t *out_xy) {
    affine pb; memcpy(&pb, b_xy, 64);
    if (!affine_is_identity(pb)) { pb.x = fe_to_mont<P>(pb.x); pb.y = fe_to_mont<P>(pb.y); }
    xyzz b = xyzz_from_affine<P>(pb);
    xyzz_double<P>(b); xyzz_add_mixed<P>(b, pb);         // 3 * pb: a genuinely projective operand (zz != 1)
    uint32_t kk[8]; memcpy(kk, k, 32);
    xyzz r = quad ? xyzz_scalar_mul_glv_q<P>(b, kk) : xyzz_scalar_mul_glv<P>(b, kk);
    affine a = jacobian_to_affine<P>(xyzz_to_jacobian<P>(r));
    a.x = fe_from_mont<P>(a.x); a.y = fe_from_mont<P>(a.y);
    memcpy(out_xy, &a, 64);
    return 0;
}
extern "C" int emu_glv_mul3(int curve, int quad, const uint8_t *b_xy, const uint8_t *k, uint8_t *out_xy) {
    return curve == 0 ? run_glv_mul<FpParams>(quad, b_xy, k, out_xy) : run_glv_mul<FqParams>(quad, b_xy, k, out_xy);
}
// the butterfly's fused (a + b, a - b) on projective operands 3a, 5b (or degenerate pairs): out = sum || diff, affine
template <class P> static int run_addsub(const uint8_t *a_xy, const uint8_t *b_xy, uint8_t *out) {
    affine pa, pb; memcpy(&pa, a_xy, 64); memcpy(&pb, b_xy, 64);
    if (!affine_is_identity(pa)) { pa.x = fe_to_mont<P>(pa.x); pa.y = fe_to_mont<P>(pa.y); }
    if (!affine_is_identity(pb)) { pb.x = fe_to_mont<P>(pb.x); pb.y = fe_to_mont<P>(pb.y); }
    xyzz a = xyzz_from_affine<P>(pa), b = xyzz_from_affine<P>(pb);
    xyzz_double<P>(a); xyzz_add_mixed<P>(a, pa);                                   // 3 pa
    xyzz_double<P>(b); xyzz_double<P>(b); xyzz_add_mixed<P>(b, pb);                // 5 pb
    xyzz sum, diff;
    xyzz_addsub_q<P, true>(a, b, sum, diff);
    affine o[2] = {jacobian_to_affine<P>(xyzz_to_jacobian<P>(sum)), jacobian_to_affine<P>(xyzz_to_jacobian<P>(diff))};
    for (int i = 0; i < 2; i++) { o[i].x = fe_from_mont<P>(o[i].x); o[i].y = fe_from_mont<P>(o[i].y); }
    memcpy(out, o, 128);
    return 0;
}
extern "C" int emu_addsub35(int curve, const uint8_t *a_xy, const uint8_t *b_xy, uint8_t *out) {
    return curve == 0 ? run_addsub<FpParams>(a_xy, b_xy, out) : run_addsub<FqParams>(a_xy, b_xy, out);
}
