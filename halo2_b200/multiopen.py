"""Host-side mirror of the multi-point opening argument, poly::multiopen
(/root/reference/halo2_proofs/src/poly/multiopen.rs:42-275, multiopen/prover.rs:18-124, multiopen/verifier.rs:14-140), over the
C ABI: same names, argument meaning and failure behaviour as the reference.

Prover: every polynomial is a ResidentPoly; the per-set combination q_i = q_i * x_1 + poly and the folds with x_2 and x_4 are
one elementwise pass each (`h2_poly_scale_add`), the divisions by (X - point) run on the device (`h2_poly_kate_division`), f's
commitment is a fixed-base MSM over the resident generators, the evaluations at x_3 one batched reduction (`h2_poly_eval`),
and the final opening is halo2_b200.opening.create_proof.  Verifier: the commitments are combined in halo2_b200.verifier.MSM
objects (scalars on the host, a few dozen terms), the interpolation through a point set is a handful of field operations
(arithmetic.rs:376-432), and the opening is halo2_b200.verifier.verify_proof -- whose Guard ends in the one big multiexp.

"The same polynomial / commitment" is object identity, as the reference compares by pointer (prover.rs:131-135,
multiopen.rs:103-114).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import lib as _l
from . import opening
from .poly import FIELDS, Blind, Params, ResidentPoly, eval_polynomial_resident, kate_division_resident
from .verifier import MSM, Guard, VerifyError
from .verifier import verify_proof as _commitment_verify_proof


class ProverQuery:
    """multiopen.rs:42-50: `poly` (a ResidentPoly of params.n coefficients, blind `blind`) is opened at `point`."""

    def __init__(self, point: int, poly: ResidentPoly, blind: Blind):
        self.point, self.poly, self.blind = int(point), poly, blind

    def _key(self):
        return id(self.poly)

    def _ref(self):
        return (self.poly, self.blind)

    def _eval(self):
        return ()                                                # Query::Eval = () for the prover (prover.rs:146-148)


class VerifierQuery:
    """multiopen.rs:53-88: a commitment (an affine point, or an MSM of commitments) claimed to open to `eval` at `point`."""

    def __init__(self, commitment, point: int, eval: int):
        self.commitment, self.point, self.eval = commitment, int(point), int(eval)

    @classmethod
    def new_commitment(cls, commitment, point: int, eval: int) -> "VerifierQuery":
        return cls(commitment, point, eval)

    @classmethod
    def new_msm(cls, msm: MSM, point: int, eval: int) -> "VerifierQuery":
        return cls(msm, point, eval)

    def _key(self):
        return id(self.commitment)

    def _ref(self):
        return self.commitment

    def _eval(self):
        return self.eval


class _CommitmentData:                                           # multiopen.rs:117-133
    def __init__(self, commitment):
        self.commitment, self.set_index, self.point_indices, self.evals = commitment, 0, [], []


def construct_intermediate_sets(queries) -> Optional[Tuple[List[_CommitmentData], List[List[int]]]]:
    """multiopen.rs:144-275: groups the queries by commitment, orders the points by first appearance, gives every distinct SET
    of points an index (by first appearance among the commitments), and returns (commitment data in first-seen order, the points
    of every set in point-index order) -- or None when a (commitment, point) pair occurs twice (:243-249)."""
    queries = list(queries)
    data: Dict[int, _CommitmentData] = {}
    point_index: Dict[int, int] = {}
    for q in queries:                                            # :162-173
        pi = point_index.setdefault(q.point, len(point_index))
        data.setdefault(q._key(), _CommitmentData(q._ref())).point_indices.append(pi)
    point_of = {i: p for p, i in point_index.items()}
    set_index: Dict[Tuple[int, ...], int] = {}
    sets_of: Dict[int, Tuple[int, ...]] = {}
    for key, d in data.items():                                  # :186-203
        ps = tuple(sorted(set(d.point_indices)))
        sets_of[key] = ps
        set_index.setdefault(ps, len(set_index))
        d.evals = [None] * len(ps)
    for q in queries:                                            # :206-250
        d, ps = data[q._key()], sets_of[q._key()]
        d.set_index = set_index[ps]
        slot = ps.index(point_index[q.point])
        if d.evals[slot] is not None:
            return None
        d.evals[slot] = q._eval()
    point_sets: List[List[int]] = [[] for _ in set_index]
    for ps, si in set_index.items():                             # :266-272
        point_sets[si] = [point_of[i] for i in ps]
    return list(data.values()), point_sets


def lagrange_interpolate(points: Sequence[int], evals: Sequence[int], modulus: int) -> List[int]:
    """arithmetic.rs:376-432: coefficients of the polynomial of degree < len(points) through (points[i], evals[i]).  Point sets
    hold a handful of rotations: host arithmetic, like the challenges."""
    assert len(points) == len(evals)
    m = modulus
    if len(points) == 1:
        return [evals[0] % m]
    out = [0] * len(points)
    for j, (x_j, e) in enumerate(zip(points, evals)):
        basis = [1]                                              # prod_{k != j} (X - x_k) / (x_j - x_k), built factor by factor
        for kk, x_k in enumerate(points):
            if kk != j:
                d = pow((x_j - x_k) % m, -1, m)
                basis = [((basis[i] if i < len(basis) else 0) * (-d * x_k) + (basis[i - 1] if i else 0) * d) % m for i in range(len(basis) + 1)]
        for i, b in enumerate(basis):
            out[i] = (out[i] + b * e) % m
    return out


def _eval_host(poly: Sequence[int], x: int, m: int) -> int:
    acc = 0
    for c in reversed(list(poly)):
        acc = (acc * x + c) % m
    return acc


def create_proof(params: Params, rng, transcript, queries) -> None:
    """multiopen::create_proof (prover.rs:18-124).  Raises ValueError for a repeated (polynomial, point) query like the
    reference's io::Error (:41-46).  `rng`: see halo2_b200.opening (f's blind is one more scalar(), drawn first, :98)."""
    n = params.n
    field = _l.SCALAR_FIELD[params.curve]
    m = FIELDS[field]
    x_1 = transcript.squeeze_challenge()                         # :38
    x_2 = transcript.squeeze_challenge()                         # :39
    sets = construct_intermediate_sets(queries)
    if sets is None:
        raise ValueError("queries iterator contains mismatching evaluations")
    poly_map, point_sets = sets
    q_polys: List[Optional[ResidentPoly]] = [None] * len(point_sets)
    q_blinds = [0] * len(point_sets)
    tmp: List[ResidentPoly] = []
    try:
        for d in poly_map:                                       # :53-73: q_i = q_i * x_1 + poly; the blinds alike
            poly, blind = d.commitment
            assert poly.len == n, "multiopen: polynomial length != params.n"
            if q_polys[d.set_index] is None:
                q_polys[d.set_index] = ResidentPoly(field, n).copy_from(poly, n)
                tmp.append(q_polys[d.set_index])
            else:
                opening._scale_add(q_polys[d.set_index], x_1, poly, 1, n)
            q_blinds[d.set_index] = (q_blinds[d.set_index] * x_1 + blind.value) % m
        q_prime: Optional[ResidentPoly] = None
        for points, q in zip(point_sets, q_polys):               # :75-96: divide by every (X - point) of the set, fold with x_2
            cur = q
            for pt in points:
                nxt = ResidentPoly(field, n)                     # zero-filled: the quotient's n - 1 coefficients and a zero on top (:84)
                tmp.append(nxt)
                kate_division_resident([cur], [pt], dst=[nxt], n=n)
                cur = nxt
            if q_prime is None:
                q_prime = cur                                    # every set has a point: cur is a quotient buffer of ours, never q_i
            else:
                opening._scale_add(q_prime, x_2, cur, 1, n)
        q_prime_blind = rng.scalar() % m                         # :98
        transcript.write_point(params.commit_resident_affine([q_prime], [Blind(q_prime_blind)])[0])   # :99-101
        x_3 = transcript.squeeze_challenge()                     # :103
        for e in eval_polynomial_resident(q_polys, [x_3] * len(q_polys), n=n):   # :107-109
            transcript.write_scalar(e)
        x_4 = transcript.squeeze_challenge()                     # :111
        p_blind = q_prime_blind
        for q, b in zip(q_polys, q_blinds):                      # :113-121  p = p * x_4 + q_i  (in q' 's buffer)
            opening._scale_add(q_prime, x_4, q, 1, n)
            p_blind = (p_blind * x_4 + b) % m
        opening.create_proof(params, rng, transcript, q_prime, Blind(p_blind), x_3)   # :123
    finally:
        for t in tmp:
            t.close()


def verify_proof(params: Params, transcript, queries, msm: MSM) -> Guard:
    """multiopen::verify_proof (verifier.rs:14-140): `msm` is the (usually empty) MSM the commitment being opened is accumulated
    into.  Raises VerifyError where the reference returns Error::OpeningError / Error::SamplingError."""
    r = msm.r
    x_1 = transcript.squeeze_challenge()                         # :31
    x_2 = transcript.squeeze_challenge()                         # :35
    sets = construct_intermediate_sets(queries)
    if sets is None:
        raise VerifyError("OpeningError")                        # :37-38
    commitment_map, point_sets = sets
    q_commitments = [MSM(params) for _ in point_sets]            # :42-45, with the next power of x_1 per set
    powers = [1] * len(point_sets)
    q_eval_sets = [[0] * len(ps) for ps in point_sets]
    scratch: List[MSM] = list(q_commitments)
    try:
        for d in reversed(commitment_map):                       # :75-83: increasing powers of x_1 from the last commitment
            s = d.set_index
            if isinstance(d.commitment, MSM):                    # :62-66
                scaled = d.commitment.clone()
                scratch.append(scaled)
                scaled.scale(powers[s])
                q_commitments[s].add_msm(scaled)
            else:
                q_commitments[s].append_term(powers[s], d.commitment)   # :59-61
            for i, e in enumerate(d.evals):                      # :68-70
                q_eval_sets[s][i] = (q_eval_sets[s][i] + e * powers[s]) % r
            powers[s] = powers[s] * x_1 % r
        try:
            q_prime_commitment = transcript.read_point()         # :87
        except Exception as e:
            raise VerifyError("SamplingError") from e
        x_3 = transcript.squeeze_challenge()                     # :91
        u = []
        for _ in q_eval_sets:                                    # :95-98
            try:
                u.append(transcript.read_scalar())
            except Exception as e:
                raise VerifyError("SamplingError") from e
        msm_eval = 0
        for points, evals, proof_eval in zip(point_sets, q_eval_sets, u):   # :102-117
            r_eval = _eval_host(lagrange_interpolate(points, evals, r), x_3, r)
            e = (proof_eval - r_eval) % r
            for pt in points:
                e = e * pow((x_3 - pt) % r, -1, r) % r
            msm_eval = (msm_eval * x_2 + e) % r
        x_4 = transcript.squeeze_challenge()                     # :121
        msm.append_term(1, q_prime_commitment)                   # :124
        v = msm_eval
        for qc, q_eval in zip(q_commitments, u):                 # :125-133
            msm.scale(x_4)
            msm.add_msm(qc)
            v = (v * x_4 + q_eval) % r
    finally:
        for t in scratch:
            t.close()
    return _commitment_verify_proof(params, msm, transcript, x_3, v)   # :136
