"""GPU parity: the permutation argument's grand product (plonk/permutation/prover.rs:98-157) composed from the device pieces --
Ast programs in the Lagrange basis for the denominators and numerators, batch_invert, the running product -- against the
reference's loops restated with big integers."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import cref, pasta  # noqa: E402

SEED = 0x48414C4F32


@pytest.mark.parametrize("field,k", [("fp", 8), ("fq", 11), ("fp", 1)])
def test_permutation_grand_product(field, k):
    import halo2_b200 as h2
    from halo2_b200.evaluator import Ast
    m = pasta.FIELDS[field]
    n = 1 << k
    d = h2.EvaluationDomain(field, 2, k, pasta.zeta_candidates(field)[0])
    cols = 3
    values = [pasta.gen_scalars(field, SEED + 1200 + j, n) for j in range(cols)]
    perms = [pasta.gen_scalars(field, SEED + 1210 + j, n) for j in range(cols)]
    beta, gamma, last_z = pasta.gen_scalars(field, SEED + 1220, 3)
    delta = 7
    # ---- the reference's loops (permutation/prover.rs:98-157)
    mv = [1] * n
    for v, s in zip(values, perms):                                  # :101-116
        mv = [x * ((beta * s_i + gamma + v_i) % m) % m for x, s_i, v_i in zip(mv, s, v)]
    mv = [pasta.inv(x, m) if x else 0 for x in mv]                   # :120 batch_invert
    deltaomega = 1
    for v in values:                                                 # :124-143
        cur = deltaomega
        for i in range(n):
            mv[i] = mv[i] * ((cur * beta + gamma + v[i]) % m) % m
            cur = cur * d.omega % m
        deltaomega = deltaomega * delta % m
    z = [last_z]
    for row in range(1, n):                                          # :150-156
        z.append(z[row - 1] * mv[row - 1] % m)
    # ---- the same on the device, nothing but the inputs going up and z coming down
    ev = h2.Evaluator(d, "lagrange")
    v_l = [ev.register_poly(cref.ints_to_bytes(v)) for v in values]
    s_l = [ev.register_poly(cref.ints_to_bytes(s)) for s in perms]
    den = None
    for v, s in zip(v_l, s_l):
        term = s * beta + Ast.constant_term(gamma) + v
        den = term if den is None else den * term
    inv_den = h2.batch_invert_resident(ev.evaluate(den))
    assert cref.bytes_to_ints(inv_den.download()) == [pasta.inv(x, m) if x else 0 for x in pasta.ast_evaluate(
        pasta.EvaluationDomain(field, 2, k, d.g_coset), "lagrange", _tuple(den), values + perms)]
    num = ev.register_poly(inv_den)
    for j, v in enumerate(v_l):
        num = num * (Ast.linear_term(pow(delta, j, m) * beta % m) + Ast.constant_term(gamma) + v)
    mv_dev = ev.evaluate(num)
    assert cref.bytes_to_ints(mv_dev.download()) == mv
    z_dev = h2.running_product_resident(mv_dev, init=last_z)
    assert cref.bytes_to_ints(z_dev.download()) == z
    # misuse
    from halo2_b200 import lib as L
    with pytest.raises(L.H2Error):
        h2.running_product_resident(mv_dev, init=1, dst=mv_dev)
    for r in (inv_den, mv_dev, z_dev):
        r.close()
    ev.close()


def _tuple(node):
    k, a = node.kind, node.args
    if k == "poly":
        return ("poly", a[0], a[1])
    if k in ("add", "mul"):
        return (k, _tuple(a[0]), _tuple(a[1]))
    if k == "scale":
        return ("scale", _tuple(a[0]), a[1])
    if k == "dp":
        return ("dp", [_tuple(t) for t in a[0]], a[1])
    return (k, a[0])
