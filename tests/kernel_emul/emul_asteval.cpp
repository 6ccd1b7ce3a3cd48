// TEST-ONLY serial execution of the expression evaluator (asteval.cuh) on the host.
#include <cstring>
#include <vector>
#include "asteval.cuh"
#include "ntt.cuh"
using namespace h2;

// polys: n_polys x n canonical; code: n_code x 4 uint32 (op, arg, shift, pad); consts canonical; omega / lin_base canonical
template <class P>
static int run_ast(const uint8_t *polys, uint32_t n_polys, uint32_t log_n, const uint32_t *code, uint32_t n_code, const uint8_t *consts, uint32_t n_consts,
                   const uint8_t *omega, const uint8_t *lin_base, uint8_t *out) {
    const uint64_t n = 1ull << log_n;
    std::vector<std::vector<fe>> p(n_polys, std::vector<fe>(n));
    std::vector<const fe *> pp(n_polys);
    for (uint32_t b = 0; b < n_polys; b++) {
        for (uint64_t i = 0; i < n; i++) { memcpy(p[b][i].v, polys + 32 * (b * n + i), 32); p[b][i] = fe_to_mont<P>(p[b][i]); }
        pp[b] = p[b].data();
    }
    std::vector<fe> cs(n_consts ? n_consts : 1), tw(n / 2 ? n / 2 : 1), pow2(64), o(n);
    for (uint32_t c = 0; c < n_consts; c++) { memcpy(cs[c].v, consts + 32 * c, 32); cs[c] = fe_to_mont<P>(cs[c]); }
    fe w; memcpy(w.v, omega, 32); w = fe_to_mont<P>(w);
    TwiddleGen<P>::pow2_body(pow2.data(), w, log_n ? log_n : 1);
    for (uint64_t t = 0; t * 32 < (n / 2 ? n / 2 : 1); t++) TwiddleGen<P>::fill_body(tw.data(), pow2.data(), n / 2 ? n / 2 : 1, t);
    AstArgs A;
    A.polys = pp.data(); A.code = reinterpret_cast<const AstInstr *>(code); A.n_code = n_code; A.consts = cs.data(); A.tw = tw.data();
    memcpy(A.lin_base.v, lin_base, 32); A.lin_base = fe_to_mont<P>(A.lin_base);
    A.log_n = log_n; A.out = o.data();
    for (uint64_t i = 0; i < n; i++) AstEval<P>::body(A, i);
    for (uint64_t i = 0; i < n; i++) { fe r = fe_from_mont<P>(o[i]); memcpy(out + 32 * i, r.v, 32); }
    return 0;
}
extern "C" int emu_ast_eval(int field, const uint8_t *polys, uint32_t n_polys, uint32_t log_n, const uint32_t *code, uint32_t n_code, const uint8_t *consts,
                            uint32_t n_consts, const uint8_t *omega, const uint8_t *lin_base, uint8_t *out) {
    if (field == 0) return run_ast<FpParams>(polys, n_polys, log_n, code, n_code, consts, n_consts, omega, lin_base, out);
    return run_ast<FqParams>(polys, n_polys, log_n, code, n_code, consts, n_consts, omega, lin_base, out);
}
