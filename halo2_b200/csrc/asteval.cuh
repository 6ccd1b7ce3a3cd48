// K15: the polynomial-expression evaluator of the quotient pipeline -- SURVEY.md section 8(f) row 3, the consumer between
// coeff_to_extended and divide_by_vanishing_poly.
//
// Restates Evaluator::evaluate (/root/reference/halo2_proofs/src/poly/evaluator.rs:129-228): an `Ast` over registered
// polynomials of one basis is applied element by element --
//   Poly(leaf)              the polynomial rotated by leaf.rotation (:151-157; rotate_left / rotate_right of the value vector,
//                           poly.rs:237-290, by `rotation` rows = rotation * 2^(extended_k - k) positions in the extended basis,
//                           poly/domain.rs:286-295)
//   Add / Mul / Scale       elementwise (:158-181)
//   DistributePowers        fold acc = acc * base + term from zero (:182-193)
//   LinearTerm(s)           s * omega^i in the Lagrange basis, s * zeta * extended_omega^i in the extended one (:538-555, :584-604)
//   ConstantTerm(s)         s everywhere (:529-536, :575-582)
// The reference walks the tree once per chunk and allocates a vector per node; here the host flattens the tree into a
// postfix program once and every thread runs it on its element with a small operand stack, reading the leaves straight from
// the resident polynomials (rotation = an index offset) and writing one output element: one launch, every operand read
// once per use, nothing intermediate in memory.
#pragma once
#include "field.cuh"

namespace h2 {

enum : uint32_t { AST_POLY = 0, AST_CONST = 1, AST_LINEAR = 2, AST_ADD = 3, AST_MUL = 4, AST_SCALE = 5, AST_NEG = 6 };
#define H2_AST_STACK 24

struct AstInstr { uint32_t op; uint32_t arg; int32_t shift; uint32_t pad; };   // POLY: arg = polynomial, shift = positions (already
                                                                                // scaled by the basis' rotation stride); CONST / LINEAR /
                                                                                // SCALE: arg = constant index
struct AstArgs {
    const fe *const *polys;     // device pointers, Montgomery form, n elements each
    const AstInstr *code;
    uint32_t n_code;
    const fe *consts;           // Montgomery form
    const fe *tw;               // omega^i, i < n / 2 (omega: the basis' root of order n); nullptr when the program has no LINEAR
    fe lin_base;                // 1 (Lagrange) or zeta (extended), Montgomery
    uint32_t log_n;
    fe *out;
};

template <class P> struct AstEval {
    static H2_HD void body(const AstArgs &A, uint64_t i) {
        const uint64_t n = 1ull << A.log_n, mask = n - 1;
        if (i >= n) return;
        fe st[H2_AST_STACK];
        uint32_t sp = 0;
        for (uint32_t pc = 0; pc < A.n_code; pc++) {
            const AstInstr in = A.code[pc];
            switch (in.op) {
            case AST_POLY: st[sp++] = fe_load(A.polys[in.arg] + ((i + (uint64_t)(int64_t)in.shift) & mask)); break;   // rotate_left by shift
            case AST_CONST: st[sp++] = fe_load(A.consts + in.arg); break;
            case AST_LINEAR: {
                const uint64_t h = n >> 1;
                fe w = (n == 1) ? fe_one<P>() : fe_load(A.tw + (i & (h - 1)));
                if (n > 1 && i >= h) w = fe_neg<P>(w);                     // omega^(n/2) = -1
                st[sp++] = fe_mul<P>(fe_mul<P>(w, A.lin_base), fe_load(A.consts + in.arg));
                break;
            }
            case AST_ADD: sp--; st[sp - 1] = fe_add<P>(st[sp - 1], st[sp]); break;
            case AST_MUL: sp--; st[sp - 1] = fe_mul<P>(st[sp - 1], st[sp]); break;
            case AST_SCALE: st[sp - 1] = fe_mul<P>(st[sp - 1], fe_load(A.consts + in.arg)); break;
            default: st[sp - 1] = fe_neg<P>(st[sp - 1]); break;
            }
        }
        fe_store(A.out + i, st[0]);
    }
};

#if defined(__CUDACC__)
template <class P> __global__ void __launch_bounds__(128) ast_eval_kernel(const AstArgs A) {
    AstEval<P>::body(A, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
}
#endif

}  // namespace h2
