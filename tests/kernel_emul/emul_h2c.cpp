// TEST-ONLY serial execution of the hash_to_curve body (h2c.cuh) on the host.
#include <cstring>
#include <vector>
#include "h2c.cuh"
using namespace h2;

template <class P> static int run_h2c(const char *domain, const uint8_t *msgs, uint32_t msg_len, int gen_params, uint64_t first, uint64_t n,
                                      uint8_t *out_xy) {
    const H2cConst K = make_h2c_const<P>(domain);
    if (!K.ok) return 1;
    std::vector<affine> o(n ? n : 1);
    for (uint64_t i = 0; i < n; i++) h2c_body<P>(msgs, msg_len, gen_params, first, K, o.data(), 0, n, i);
    memcpy(out_xy, o.data(), n * 64);
    return 0;
}
extern "C" int emu_hash_to_curve(int curve, const char *domain, const uint8_t *msgs, uint32_t msg_len, int gen_params, uint64_t first,
                                 uint64_t n, uint8_t *out_xy) {
    return curve == 0 ? run_h2c<FpParams>(domain, msgs, msg_len, gen_params, first, n, out_xy)
                      : run_h2c<FqParams>(domain, msgs, msg_len, gen_params, first, n, out_xy);
}
// BLAKE2b-512 of one buffer (unkeyed): checks the hash core on its own against hashlib
extern "C" void emu_blake2b(const uint8_t *in, uint32_t len, uint8_t *out64) {
    Blake2b S;
    uint8_t d[64];
    b2_init(S);
    b2_update(S, in, len);
    b2_final(S, d);
    memcpy(out64, d, 64);
}
