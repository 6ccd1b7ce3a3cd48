"""Poseidon P128Pow5T3 permutation restated over an abstract (add, mul, pow5) field backend.

Follows /root/reference/halo2_poseidon/src/lib.rs:106-150 (R_F/2 full rounds, R_P partial
rounds, R_F/2 full rounds; full round = add rc, x^5 on every word, MDS; partial round = add rc,
x^5 on word 0 only, MDS).  Used as a field-arithmetic known-answer test: it exercises modular
add and mul with the reference's own constants and expected outputs
(halo2_poseidon/src/test_vectors.rs, checked there by p128pow5t3.rs:258-290).
"""


def permute(state, rc, mds, add, mul, pow5, r_f=8, r_p=56):
    T = 3
    rcs = [rc[3 * i:3 * i + 3] for i in range(r_f + r_p)]
    m = [mds[3 * i:3 * i + 3] for i in range(T)]

    def apply_mds(st):
        out = []
        for i in range(T):
            acc = 0
            for j in range(T):
                acc = add(acc, mul(m[i][j], st[j]))
            out.append(acc)
        return out

    rounds = ["f"] * (r_f // 2) + ["p"] * r_p + ["f"] * (r_f // 2)
    st = list(state)
    for kind, c in zip(rounds, rcs):
        st = [add(w, k) for w, k in zip(st, c)]
        if kind == "f":
            st = [pow5(w) for w in st]
        else:
            st[0] = pow5(st[0])
        st = apply_mds(st)
    return st
