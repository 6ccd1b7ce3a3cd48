// TEST-ONLY serial execution of the point (de)compression bodies (codec.cuh) on the host.
#include <cstring>
#include <vector>
#include "codec.cuh"
using namespace h2;

template <class P> static int run_compress(const uint8_t *in_xy, uint64_t n, uint8_t *out) {
    std::vector<affine> a(n ? n : 1);
    std::vector<fe> o(n ? n : 1);
    memcpy(a.data(), in_xy, n * 64);
    for (uint64_t i = 0; i < n; i++) Codec<P>::compress_body(a.data(), 0, o.data(), n, i);
    memcpy(out, o.data(), n * 32);
    return 0;
}
template <class P> static uint32_t run_decompress(const uint8_t *in, uint64_t n, uint8_t *out_xy) {
    std::vector<fe> a(n ? n : 1);
    std::vector<affine> o(n ? n : 1);
    memcpy(a.data(), in, n * 32);
    const SqrtConst K = make_sqrt_const<P>();
    uint32_t bad = 0xffffffffu;
    for (uint64_t i = 0; i < n; i++) Codec<P>::decompress_body(a.data(), o.data(), 0, K, &bad, n, i);
    memcpy(out_xy, o.data(), n * 64);
    return bad;
}
extern "C" int emu_compress(int curve, const uint8_t *in_xy, uint64_t n, uint8_t *out) {
    return curve == 0 ? run_compress<FpParams>(in_xy, n, out) : run_compress<FqParams>(in_xy, n, out);
}
// returns the index of the first invalid encoding, 0xffffffff when all are valid
extern "C" uint32_t emu_decompress(int curve, const uint8_t *in, uint64_t n, uint8_t *out_xy) {
    return curve == 0 ? run_decompress<FpParams>(in, n, out_xy) : run_decompress<FqParams>(in, n, out_xy);
}
