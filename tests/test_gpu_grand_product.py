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


@pytest.mark.parametrize("field,k", [("fp", 9), ("fq", 6)])
def test_lookup_argument_on_device(field, k):
    """The lookup argument's prover steps between the compressed columns and the product commitment, on resident columns:
    permute_expression_pair (plonk/lookup/prover.rs:563-647, h2_poly_lookup_permute) with the caller's blinding rows copied
    in (:625-627), then commit_product's grand product (:279-337) composed like the permutation's -- an Ast program for the
    denominators (a' + beta)(s' + gamma), batch_invert, an Ast program for the numerators, the running product from 1 --
    against the reference's loops restated with big integers; the reference's own sanity identities hold (:343-378)."""
    import random
    import halo2_b200 as h2
    from halo2_b200.evaluator import Ast
    m = pasta.FIELDS[field]
    n = 1 << k
    bf = 5                                     # blinding_factors of the benches/plonk.rs-sized circuits
    u = n - (bf + 1)                           # usable rows, :572-573
    rnd = random.Random(31 + k)
    pool = [rnd.randrange(m) for _ in range(max(4, n // 8))]
    table = (pool + [rnd.choice(pool) for _ in range(n)])[:n]
    inputs = [rnd.choice(table[:u]) for _ in range(n)]
    beta, gamma = pasta.gen_scalars(field, SEED + 1300, 2)
    tail_a = [rnd.randrange(m) for _ in range(n)]          # the rows from u on are the random blinding values
    tail_s = [rnd.randrange(m) for _ in range(n)]
    # ---- the reference's steps
    pa, ps = pasta.permute_expression_pair(field, inputs, table, u)
    pa, ps = pa + tail_a[u:], ps + tail_s[u:]
    lp = [(beta + x) * (gamma + y) % m for x, y in zip(pa, ps)]                      # :281-291
    lp = [pasta.inv(x, m) if x else 0 for x in lp]                                   # :295
    lp = [x * ((a + beta) % m) % m * ((s + gamma) % m) % m for x, a, s in zip(lp, inputs, table)]   # :300-310
    z = [1]
    for cur in lp:                                                                   # :327-337 (scan from ONE, n - bf rows kept)
        z.append(z[-1] * cur % m)
    z = z[:n - bf]
    assert z[u] == 1                                                                 # :378
    for i in range(u):                                                               # :351-372
        assert z[i + 1] * ((beta + pa[i]) % m) % m * ((gamma + ps[i]) % m) % m == z[i] * ((inputs[i] + beta) % m) % m * ((table[i] + gamma) % m) % m
    # ---- the same on the device
    d = h2.EvaluationDomain(field, 2, k, pasta.zeta_candidates(field)[0])
    a_l, s_l = h2.ResidentPoly(field, n, cref.ints_to_bytes(inputs)), h2.ResidentPoly(field, n, cref.ints_to_bytes(table))
    pa_l, ps_l = h2.permute_expression_pair_resident(a_l, s_l, u)
    ta, ts = h2.ResidentPoly(field, n, cref.ints_to_bytes(tail_a)), h2.ResidentPoly(field, n, cref.ints_to_bytes(tail_s))
    pa_l.copy_from(ta, n - u, src_off=u, dst_off=u)
    ps_l.copy_from(ts, n - u, src_off=u, dst_off=u)
    assert cref.bytes_to_ints(pa_l.download()) == pa and cref.bytes_to_ints(ps_l.download()) == ps
    ev = h2.Evaluator(d, "lagrange")
    A, S, PA, PS = (ev.register_poly(p) for p in (a_l, s_l, pa_l, ps_l))
    den = (PA + Ast.constant_term(beta)) * (PS + Ast.constant_term(gamma))
    inv_den = h2.batch_invert_resident(ev.evaluate(den))
    num = ev.register_poly(inv_den) * (A + Ast.constant_term(beta)) * (S + Ast.constant_term(gamma))
    lp_dev = ev.evaluate(num)
    assert cref.bytes_to_ints(lp_dev.download()) == lp
    z_dev = h2.running_product_resident(lp_dev, init=1)
    assert cref.bytes_to_ints(z_dev.download())[:n - bf] == z
    for r in (a_l, s_l, pa_l, ps_l, ta, ts, inv_den, lp_dev, z_dev):
        r.close()
    ev.close()
