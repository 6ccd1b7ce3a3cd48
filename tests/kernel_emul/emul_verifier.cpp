// TEST-ONLY serial execution of the verifier's kernels (verifier.cuh) with the launch shapes of capi_poly.cu.
#include <cstring>
#include <vector>
#include "verifier.cuh"
using namespace h2;

template <class P> static fe load_mont(const uint8_t *b) { fe x; memcpy(x.v, b, 32); return fe_to_mont<P>(x); }

// dst (2^k canonical, read when accumulate) (+)= compute_s(u, init); u: k canonical challenges
template <class P> static int run_compute_s(const uint8_t *u_in, uint32_t k, const uint8_t *init, int accumulate, uint8_t *dst_io) {
    const uint64_t n = 1ull << k;
    std::vector<fe> u(k), d(n);
    for (uint32_t j = 0; j < k; j++) u[j] = load_mont<P>(u_in + 32 * j);
    for (uint64_t i = 0; i < n; i++) d[i] = load_mont<P>(dst_io + 32 * i);
    const fe in0 = load_mont<P>(init);
    const uint64_t groups = 1ull << (k - (k < 2 ? k : 2));
    const uint64_t threads = (groups + 127) / 128 * 128;        // the grid capi_poly.cu launches, idle threads included
    for (uint64_t t = 0; t < threads; t++) VerifierOps<P>::compute_s_body(d.data(), u.data(), k, in0, accumulate, t);
    for (uint64_t i = 0; i < n; i++) { fe r = fe_from_mont<P>(d[i]); memcpy(dst_io + 32 * i, r.v, 32); }
    return 0;
}
extern "C" int emu_compute_s(int field, const uint8_t *u, uint32_t k, const uint8_t *init, int accumulate, uint8_t *dst_io) {
    return field == 0 ? run_compute_s<FpParams>(u, k, init, accumulate, dst_io) : run_compute_s<FqParams>(u, k, init, accumulate, dst_io);
}
// dst = a * dst + b * src (src may be null)
template <class P> static int run_scale_add(uint8_t *dst_io, const uint8_t *a, const uint8_t *src, const uint8_t *b, uint64_t n) {
    std::vector<fe> d(n), s(n);
    for (uint64_t i = 0; i < n; i++) { d[i] = load_mont<P>(dst_io + 32 * i); if (src) s[i] = load_mont<P>(src + 32 * i); }
    const fe fa = load_mont<P>(a), fb = src ? load_mont<P>(b) : fe_zero();
    const uint64_t threads = (n + 255) / 256 * 256;
    for (uint64_t t = 0; t < threads; t++) VerifierOps<P>::scale_add_body(d.data(), src ? s.data() : nullptr, fa, fb, n, t);
    for (uint64_t i = 0; i < n; i++) { fe r = fe_from_mont<P>(d[i]); memcpy(dst_io + 32 * i, r.v, 32); }
    return 0;
}
extern "C" int emu_scale_add(int field, uint8_t *dst_io, const uint8_t *a, const uint8_t *src, const uint8_t *b, uint64_t n) {
    return field == 0 ? run_scale_add<FpParams>(dst_io, a, src, b, n) : run_scale_add<FqParams>(dst_io, a, src, b, n);
}
