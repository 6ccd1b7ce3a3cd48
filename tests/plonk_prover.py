"""plonk::create_proof driven by a PINNED verifying key and the circuit's columns, through the oracle (test infrastructure).

Why it exists: the prover-side functions of the path -- lagrange_to_coeff / coeff_to_extended / extended_to_coeff (best_fft at
G = scalar), divide_by_vanishing_poly, permute_expression_pair, eval_polynomial, kate_division, commit / commit_lagrange, the
multi-point opening and the opening argument -- have no reference-held input -> output vector of their own.  The VERIFIER is
pinned on the reference's sixteen golden proofs (tests/plonk_verifier.py, tests/test_golden_proofs.py); so a REAL proof of the
reference's own test circuit, produced by the oracle's restatements of those functions under the reference's own golden
verifying key, and ACCEPTED by that pinned verifier, ties every one of them to the reference's verification equation: a wrong
butterfly, a wrong zeta power, a mis-ordered fold anywhere and the final multiexp is not the identity.

What is restated, in the reference's order (halo2_proofs/src/plonk):
  prover.rs:43-727               create_proof: instance / advice commitments, theta, lookups, beta / gamma, permutation and
                                 lookup products, the vanishing argument, y, h(X), x, every evaluation, the query list
  permutation/prover.rs:42-173   commit (the grand product over chunks of columns, last_z chaining, blinding rows),
                                 :176-283 construct (the expressions), :296-343 evaluate, :346-394 open
  lookup/prover.rs:63-203        commit_permuted, :206-300 commit_product, :303-383 construct, :386-425 evaluate, :428-470 open
  vanishing/prover.rs:41-60      commit (the random polynomial), :64-118 construct (h pieces), :121-150 evaluate, :153-175 open
  keygen.rs:306-325              l_0, l_blind, l_last
h(X) is evaluated point by point over the extended coset (2^extended_k points: 128 for the k = 5 test circuit) with the same
scalar formulas the verifier applies at x -- a rotation by r rows is a shift by r * 2^(extended_k - k) coset points.
The randomness (blinding rows, blinds, the random polynomial, the opening's) comes from a seeded generator.
"""
from __future__ import annotations

from typing import List

from oracle import pasta
from tests import plonk_verifier as PV


def create_proof(c: pasta.Curve, g, g_lagrange, w, u, vk: PV.PinnedKey, fixed: List[List[int]], sigma: List[List[int]],
                 advice: List[List[List[int]]], instances: List[List[List[int]]], rng, transcript, zeta: int, delta: int) -> None:
    """`fixed`, `sigma`: the circuit's fixed columns and permutation polynomials as Lagrange values (what keygen computes);
    `advice[proof][column]`: the witness, 2^k values each (the last blinding_factors + 1 rows are overwritten with randomness,
    prover.rs:276-282); `instances[proof][column]`: public inputs.  `transcript`: write_point / write_scalar / common_* /
    squeeze_challenge on affine tuples and ints."""
    m = vk.scalar_modulus
    k, n = vk.k, 1 << vk.k
    bf = vk.blinding_factors()
    usable = n - (bf + 1)
    cs_degree = vk.degree()
    chunk_len = cs_degree - 2
    D = pasta.EvaluationDomain(c.scalar, cs_degree, k, zeta)
    assert D.extended_k == vk.extended_k and D.omega == vk.omega
    L, step = D.extended_len(), 1 << (D.extended_k - k)
    num_proofs = len(advice)
    commit_l = lambda vals, blind: pasta.to_affine(c, pasta.best_multiexp(c, list(vals) + [blind], list(g_lagrange) + [w]))
    commit_c = lambda vals, blind: pasta.to_affine(c, pasta.best_multiexp(c, list(vals) + [blind], list(g) + [w]))
    to_coeff, to_ext = D.lagrange_to_coeff, D.coeff_to_extended
    evalp = lambda poly, x: pasta.eval_polynomial_mod(m, poly, x)

    transcript.common_scalar(vk.transcript_repr())                 # prover.rs:56  vk.hash_into
    # ---- instance columns, :73-126 ----
    inst_vals, inst_polys, inst_cosets = [], [], []
    for inst in instances:
        vals = []
        for col in inst:
            assert len(col) <= usable
            vals.append([v % m for v in col] + [0] * (n - len(col)))
        for v in vals:
            transcript.common_point(commit_l(v, 1))               # Blind::default()
        polys = [to_coeff(v) for v in vals]
        inst_vals.append(vals), inst_polys.append(polys), inst_cosets.append([to_ext(p) for p in polys])
    # ---- advice columns, :135-321 ----
    adv_vals, adv_polys, adv_cosets, adv_blinds = [], [], [], []
    for cols in advice:
        vals = [[v % m for v in col[:usable]] + [rng.scalar() for _ in range(n - usable)] for col in cols]     # :276-282
        blinds = [rng.scalar() for _ in vals]
        for v, b in zip(vals, blinds):
            transcript.write_point(commit_l(v, b))
        polys = [to_coeff(v) for v in vals]
        adv_vals.append(vals), adv_polys.append(polys), adv_cosets.append([to_ext(p) for p in polys]), adv_blinds.append(blinds)
    fixed_polys = [to_coeff(f) for f in fixed]
    fixed_cosets = [to_ext(p) for p in fixed_polys]
    sigma_polys = [to_coeff(s) for s in sigma]
    sigma_cosets = [to_ext(p) for p in sigma_polys]
    ind = lambda rows: to_ext(to_coeff([1 if r in rows else 0 for r in range(n)]))
    l0, l_blind, l_last = ind({0}), ind(set(range(n - bf, n))), ind({n - bf - 1})          # keygen.rs:306-325

    def rows_eval(expr, pr):                                       # an Expression over the Lagrange values, row by row (rotations wrap)
        out = []
        for row in range(n):
            at = lambda cols, qs: [cols[col][(row + r) % n] for col, r in qs]
            out.append(vk.evaluate(expr, m, at(fixed, vk.fixed_queries), at(adv_vals[pr], vk.advice_queries), at(inst_vals[pr], vk.instance_queries)))
        return out

    theta = transcript.squeeze_challenge()                         # :367
    # ---- lookups: permuted columns, lookup/prover.rs:63-203 ----
    lookups = []
    for pr in range(num_proofs):
        per = []
        for inp, tab in vk.lookups:
            def compress(exprs):
                acc = [0] * n
                for e in exprs:
                    ev = rows_eval(e, pr)
                    acc = [(a * theta + b) % m for a, b in zip(acc, ev)]
                return acc
            ci, ct = compress(inp), compress(tab)
            pi, pt = pasta.permute_expression_pair(c.scalar, ci, ct, usable)                # :563-647, usable rows
            pi = list(pi) + [rng.scalar() for _ in range(bf + 1)]                           # :623-624: the blinding rows, input first
            pt = list(pt) + [rng.scalar() for _ in range(bf + 1)]
            bi = rng.scalar()                                       # commit_values, :160-170: input then table
            cmi = commit_l(pi, bi)
            bt = rng.scalar()
            cmt = commit_l(pt, bt)
            transcript.write_point(cmi)
            transcript.write_point(cmt)
            per.append({"ci": ci, "ct": ct, "pi": pi, "pt": pt, "pi_poly": to_coeff(pi), "pt_poly": to_coeff(pt), "bi": bi, "bt": bt})
        lookups.append(per)
    beta = transcript.squeeze_challenge()                          # :405
    gamma = transcript.squeeze_challenge()                         # :408
    # ---- permutation products, permutation/prover.rs:42-173 ----
    perms = []
    col_vals = lambda pr, col: {"Advice": adv_vals[pr], "Fixed": fixed, "Instance": inst_vals[pr]}[col[0]][col[1]]
    for pr in range(num_proofs):
        sets, deltaomega, last_z = [], 1, 1
        for ci in range(0, len(vk.permutation_columns), chunk_len):
            cols = vk.permutation_columns[ci:ci + chunk_len]
            mod = [1] * n
            for col, sg in zip(cols, sigma[ci:ci + chunk_len]):     # :77-94
                v = col_vals(pr, col)
                mod = [a * ((beta * s + gamma + x) % m) % m for a, s, x in zip(mod, sg, v)]
            mod = [pasta.inv(a, m) if a else 0 for a in mod]        # batch_invert, :97
            for col in cols:                                        # :101-121
                v = col_vals(pr, col)
                cur = deltaomega
                for row in range(n):
                    mod[row] = mod[row] * ((cur * beta + gamma + v[row]) % m) % m
                    cur = cur * D.omega % m
                deltaomega = deltaomega * delta % m
            z = [last_z]
            for row in range(1, n):                                 # :126-133
                z.append(z[row - 1] * mod[row - 1] % m)
            for row in range(n - bf, n):                            # :136-138
                z[row] = rng.scalar()
            last_z = z[n - (bf + 1)]                                # :140
            blind = rng.scalar()
            transcript.write_point(commit_l(z, blind))              # :144-157
            zp = to_coeff(z)
            sets.append({"poly": zp, "coset": to_ext(zp), "blind": blind})
        perms.append(sets)
    # ---- lookup products, lookup/prover.rs:206-300 ----
    for pr in range(num_proofs):
        for lk in lookups[pr]:
            prod = [(beta + a) * (gamma + s) % m for a, s in zip(lk["pi"], lk["pt"])]
            prod = [pasta.inv(p, m) if p else 0 for p in prod]
            prod = [p * ((a + beta) % m) % m * ((s + gamma) % m) % m for p, a, s in zip(prod, lk["ci"], lk["ct"])]
            z, state = [], 1
            for cur in [1] + prod:                                  # :257-263: scan, take n - bf, then bf random values
                state = state * cur % m
                z.append(state)
            z = z[:n - bf] + [rng.scalar() for _ in range(bf)]
            assert z[0] == 1 and z[usable] == 1                     # the reference's sanity checks, :270, :290
            lk["zb"] = rng.scalar()
            transcript.write_point(commit_l(z, lk["zb"]))
            lk["z_poly"] = to_coeff(z)
    # ---- vanishing argument: the random polynomial, vanishing/prover.rs:41-60 ----
    random_poly = rng.poly(n)
    random_blind = rng.scalar()
    transcript.write_point(commit_c(random_poly, random_blind))
    y = transcript.squeeze_challenge()                             # :458
    # ---- h(X) over the extended coset: gates, permutation, lookups per proof, folded by y (prover.rs:460-564) ----
    xs, cur = [], D.g_coset
    for _ in range(L):
        xs.append(cur)
        cur = cur * D.extended_omega % m
    rot = lambda arr, i, r: arr[(i + r * step) % L]
    lk_cosets = [[{kk: to_ext(lk[kk + "_poly"]) for kk in ("pi", "pt", "z")} for lk in per] for per in lookups]
    lk_compressed = [[{kk: to_ext(to_coeff(lk[kk])) for kk in ("ci", "ct")} for lk in per] for per in lookups]
    last_rot = -(bf + 1)
    num = []
    for i in range(L):
        acc = 0
        active = (1 - (l_last[i] + l_blind[i])) % m
        for pr in range(num_proofs):
            at = lambda cosets, qs: [rot(cosets[col], i, r) for col, r in qs]
            fe, ae, ie = at(fixed_cosets, vk.fixed_queries), at(adv_cosets[pr], vk.advice_queries), at(inst_cosets[pr], vk.instance_queries)
            exprs = [vk.evaluate(gate, m, fe, ae, ie) for gate in vk.gates]
            sets = perms[pr]                                        # permutation/prover.rs:200-283
            if sets:
                zc = [s["coset"] for s in sets]
                exprs.append((1 - zc[0][i]) * l0[i] % m)
                exprs.append((zc[-1][i] * zc[-1][i] - zc[-1][i]) * l_last[i] % m)
                for a in range(1, len(sets)):
                    exprs.append((zc[a][i] - rot(zc[a - 1], i, last_rot)) * l0[i] % m)
                colc = lambda col: {"Advice": adv_cosets[pr], "Fixed": fixed_cosets, "Instance": inst_cosets[pr]}[col[0]][col[1]][i]
                for a, st in enumerate(sets):
                    cols = vk.permutation_columns[a * chunk_len:(a + 1) * chunk_len]
                    left = rot(zc[a], i, 1)
                    for col, sc in zip(cols, sigma_cosets[a * chunk_len:(a + 1) * chunk_len]):
                        left = left * ((colc(col) + beta * sc[i] + gamma) % m) % m
                    right = zc[a][i]
                    cur_delta = beta * xs[i] % m * pow(delta, a * chunk_len, m) % m
                    for col in cols:
                        right = right * ((colc(col) + cur_delta + gamma) % m) % m
                        cur_delta = cur_delta * delta % m
                    exprs.append((left - right) * active % m)
            for lc, lcc in zip(lk_cosets[pr], lk_compressed[pr]):   # lookup/prover.rs:318-372
                z_, a_, s_ = lc["z"], lc["pi"], lc["pt"]
                exprs.append((1 - z_[i]) * l0[i] % m)
                exprs.append((z_[i] * z_[i] - z_[i]) * l_last[i] % m)
                left = rot(z_, i, 1) * ((a_[i] + beta) % m) % m * ((s_[i] + gamma) % m) % m
                right = z_[i] * ((lcc["ci"][i] + beta) % m) % m * ((lcc["ct"][i] + gamma) % m) % m
                exprs.append((left - right) * active % m)
                exprs.append((a_[i] - s_[i]) * l0[i] % m)
                exprs.append((a_[i] - s_[i]) * (a_[i] - rot(a_, i, -1)) % m * active % m)
            for e in exprs:                                          # Ast::distribute_powers, vanishing/prover.rs:78
                acc = (acc * y + e) % m
        num.append(acc)
    h = D.extended_to_coeff(D.divide_by_vanishing_poly(num))       # vanishing/prover.rs:85-88
    assert len(h) == n * (cs_degree - 1)
    h_pieces = [h[a * n:(a + 1) * n] for a in range(cs_degree - 1)]
    h_blinds = [rng.scalar() for _ in h_pieces]
    for piece, b in zip(h_pieces, h_blinds):
        transcript.write_point(commit_c(piece, b))
    x = transcript.squeeze_challenge()                             # :566
    xn = pow(x, n, m)
    rotx = lambda r: D.rotate_omega(x, r)
    # ---- evaluations, :569-640 ----
    for pr in range(num_proofs):
        for col, r in vk.instance_queries:
            transcript.write_scalar(evalp(inst_polys[pr][col], rotx(r)))
    for pr in range(num_proofs):
        for col, r in vk.advice_queries:
            transcript.write_scalar(evalp(adv_polys[pr][col], rotx(r)))
    for col, r in vk.fixed_queries:
        transcript.write_scalar(evalp(fixed_polys[col], rotx(r)))
    h_poly, h_blind = [0] * n, 0                                   # vanishing/prover.rs:128-138
    for piece, b in zip(reversed(h_pieces), reversed(h_blinds)):
        h_poly = [(a * xn + p) % m for a, p in zip(h_poly, piece)]
        h_blind = (h_blind * xn + b) % m
    transcript.write_scalar(evalp(random_poly, x))
    for sp in sigma_polys:                                         # pk.permutation.evaluate, permutation/prover.rs:286-294
        transcript.write_scalar(evalp(sp, x))
    for pr in range(num_proofs):                                   # permutation/prover.rs:296-343
        sets = perms[pr]
        for a, st in enumerate(sets):
            transcript.write_scalar(evalp(st["poly"], x))
            transcript.write_scalar(evalp(st["poly"], rotx(1)))
            if a + 1 < len(sets):
                transcript.write_scalar(evalp(st["poly"], rotx(last_rot)))
    for pr in range(num_proofs):                                   # lookup/prover.rs:386-425
        for lk in lookups[pr]:
            for poly, r in ((lk["z_poly"], 0), (lk["z_poly"], 1), (lk["pi_poly"], 0), (lk["pi_poly"], -1), (lk["pt_poly"], 0)):
                transcript.write_scalar(evalp(poly, rotx(r)))
    # ---- the query list, :655-724, and the multi-point opening ----
    Q = pasta.ProverQuery
    queries = []
    for pr in range(num_proofs):
        for col, r in vk.instance_queries:
            queries.append(Q(rotx(r), inst_polys[pr][col], 1))
        for col, r in vk.advice_queries:
            queries.append(Q(rotx(r), adv_polys[pr][col], adv_blinds[pr][col]))
        sets = perms[pr]                                           # permutation/prover.rs:346-394
        for st in sets:
            queries.append(Q(x, st["poly"], st["blind"]))
            queries.append(Q(rotx(1), st["poly"], st["blind"]))
        for st in list(reversed(sets))[1:]:
            queries.append(Q(rotx(last_rot), st["poly"], st["blind"]))
        for lk in lookups[pr]:                                     # lookup/prover.rs:428-470
            queries += [Q(x, lk["z_poly"], lk["zb"]), Q(x, lk["pi_poly"], lk["bi"]), Q(x, lk["pt_poly"], lk["bt"]),
                        Q(rotx(-1), lk["pi_poly"], lk["bi"]), Q(rotx(1), lk["z_poly"], lk["zb"])]
    for col, r in vk.fixed_queries:
        queries.append(Q(rotx(r), fixed_polys[col], 1))
    for sp in sigma_polys:                                         # pk.permutation.open, permutation/prover.rs:397-409
        queries.append(Q(x, sp, 1))
    queries.append(Q(x, h_poly, h_blind))                          # vanishing/prover.rs:153-175
    queries.append(Q(x, random_poly, random_blind))
    pasta.multiopen_create_proof(c, g, w, u, rng, transcript, queries)                      # :726
