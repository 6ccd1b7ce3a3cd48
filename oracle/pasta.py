"""CPU oracle (TEST INFRASTRUCTURE ONLY) -- exact big-integer restatement of the
halo2 MSM + FFT hot path over the Pasta fields/curves.

*** This file is the checker, never the product. Only tests/, __graft_entry__.smoke()
*** and bench.py's cpu_baseline / --impl reference leg may import it.

What it restates (all file:line relative to /root/reference/halo2_proofs/src):
  * best_multiexp            arithmetic.rs:143-180  (+ Bucket/Buckets :29-112)
  * small_multiexp           arithmetic.rs:116-136
  * best_fft                 arithmetic.rs:192-255  (+ recursive_butterfly_arithmetic :258-295)
  * parallelize              arithmetic.rs:345-362  (chunking only; serial here)
  * Params::commit{,_lagrange}  poly/commitment.rs:119-150
  * Params::new               poly/commitment.rs:38-114  (hash_to_curve restated from RFC 9380, see below)
  * EvaluationDomain::{new, lagrange_to_coeff, coeff_to_extended, extended_to_coeff,
    distribute_powers_zeta, ifft}   poly/domain.rs:40-146, 227-255, 303-325, 357-383
  * Evaluator::evaluate over an Ast              poly/evaluator.rs:129-228 (+ the BasisOps of :522-607)
  * point compression and Params::{write, read}   book/src/background/curves.md:203-240, poly/commitment.rs:168-205
  * the IPA round loop       poly/commitment/prover.rs:100-142, :154-166 (transcript factored out: challenges
    and randomness are inputs, the points / scalar written to the transcript are outputs)
  * permute_expression_pair  plonk/lookup/prover.rs:563-647 (the lookup argument's permuted columns, usable rows only)
  * the verifier's side: MSM (poly/commitment/msm.rs:9-178), commitment::verify_proof / Guard / compute_s / compute_b
    (poly/commitment/verifier.rs:13-171), commitment::create_proof whole (poly/commitment/prover.rs:36-151)
  * the multi-point opening argument: poly/multiopen.rs:144-275, multiopen/prover.rs:18-124, multiopen/verifier.rs:14-140,
    lagrange_interpolate arithmetic.rs:376-432

Third-party arithmetic that is NOT in the reference tree: crate `pasta_curves 0.5.1`
(Cargo.lock:1303-1306), `ff 0.13.0`, `group 0.13.0`.  Its published algorithm is restated
here from the mathematical definition: Fp/Fq are prime fields with the two moduli pinned at
halo2_proofs/tests/plonk_api.rs:591-592; Pallas/Vesta are y^2 = x^3 + 5 over Fp/Fq
(book/src/background/curves.md); identity is encoded as affine (0, 0)
(book/src/background/curves.md:226-230).

PARITY PINNING STATUS: PINNED on reference-held vectors.
  * moduli, ROOT_OF_UNITY (via the k=5 / k=11 omegas in the reference goldens), field mul/add/x^5
    (halo2_poseidon test vectors), point decompression (26 golden points);
  * best_multiexp + best_fft at G = curve point + hash_to_curve + ZETA + DELTA: all 19 golden
    commitments of the plonk_api verifying key (tests/plonk_api.rs:958-982) are reproduced from
    first principles -- Params::new(5) through hash_to_curve("Halo2-Parameters"), the EC-iFFT
    (ec_fft below), and commit_lagrange = best_multiexp over g_lagrange ++ [w] -- by this file and
    by oracle/halo2_oracle.c (tests/test_oracle_golden.py), and by the device path
    (tests/test_gpu_golden.py);
  * best_fft at G = scalar has no reference-held input->output vector of its own; it is the same
    butterfly network as the pinned G = curve-point instantiation (one generic function,
    arithmetic.rs:192-295), checked against the DFT definition for true roots of unity and, through
    the domain constants it is used with, against the pinned omegas -- and it is tied to the
    reference's verification equation end to end: a REAL proof of the reference's own test circuit
    (tests/plonk_api.rs:21-420), produced by this file's lagrange_to_coeff / coeff_to_extended /
    extended_to_coeff / divide_by_vanishing_poly / permute_expression_pair / kate_division /
    eval_polynomial / commit / multiopen / opening restatements under the reference's golden verifying
    key (tests/plonk_prover.py), is accepted by the golden-proof-pinned verifier
    (tests/test_real_proof.py); a wrong butterfly or zeta power anywhere makes it fail.
  * the verifier's side -- MSM, verify_proof / Guard, compute_s / compute_b, the multiopen verifier, lagrange_interpolate -- on
    the reference's sixteen GOLDEN PROOFS (halo2_proofs/tests/plonk_api_proof.bin and the fifteen proof_*.bin of
    halo2_gadgets/src/test_circuits/circuit_data/): every one is accepted under its pinned key, every tampered one rejected
    (tests/test_golden_proofs.py, tests/plonk_verifier.py; on the device: tests/test_gpu_golden_proofs.py).
  The reference itself cannot be built here (no Rust toolchain; pasta_curves un-vendored).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------------------
# Fields.  tests/plonk_api.rs:591-592
# --------------------------------------------------------------------------------------
P_MOD = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001  # Fp: Pallas base, Vesta scalar
Q_MOD = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001  # Fq: Vesta base, Pallas scalar
S_2ADICITY = 32  # book/src/background/fields.md:198-204
MULT_GEN = 5  # multiplicative generator used by pasta_curves for both fields

FIELDS = {"fp": P_MOD, "fq": Q_MOD}


def field_modulus(field: str) -> int:
    return FIELDS[field]


def root_of_unity(field: str) -> int:
    """ROOT_OF_UNITY = 5^T, T = (m-1) >> 32 (order exactly 2^32).  Pinned against the
    reference's k=5 and k=11 omegas in tests/test_oracle_golden.py."""
    m = FIELDS[field]
    return pow(MULT_GEN, (m - 1) >> S_2ADICITY, m)


def omega_for_k(field: str, k: int) -> int:
    """domain.rs:58-78: omega = ROOT_OF_UNITY^(2^(S-k))."""
    m = FIELDS[field]
    w = root_of_unity(field)
    for _ in range(k, S_2ADICITY):
        w = w * w % m
    return w


def zeta_candidates(field: str) -> Tuple[int, int]:
    """The two primitive cube roots of unity.  Which one pasta_curves calls ZETA is not
    pinned by any in-tree golden; the engine therefore takes zeta as an argument
    (domain.rs:85 reads F::ZETA)."""
    m = FIELDS[field]
    z = pow(MULT_GEN, (m - 1) // 3, m)
    return z, z * z % m


def inv(a: int, m: int) -> int:
    return pow(a, m - 2, m)


# --------------------------------------------------------------------------------------
# Curves: y^2 = x^3 + 5.  Pallas over Fp (scalars Fq), Vesta over Fq (scalars Fp).
# Affine point = (x, y) ints, identity = None.  Jacobian = (X, Y, Z), identity Z == 0.
# --------------------------------------------------------------------------------------
CURVE_B = 5


@dataclass(frozen=True)
class Curve:
    name: str
    base: str  # coordinate field
    scalar: str  # scalar field

    @property
    def p(self) -> int:
        return FIELDS[self.base]

    @property
    def r(self) -> int:
        return FIELDS[self.scalar]


PALLAS = Curve("pallas", "fp", "fq")
VESTA = Curve("vesta", "fq", "fp")
CURVES = {"pallas": PALLAS, "vesta": VESTA}

Affine = Optional[Tuple[int, int]]
Jac = Tuple[int, int, int]

JAC_ID: Jac = (0, 1, 0)


def on_curve(c: Curve, pt: Affine) -> bool:
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - CURVE_B) % c.p == 0


def generator(c: Curve) -> Affine:
    """(-1, 2): on both curves; the concrete point the reference's msm test uses
    (poly/commitment/msm.rs:181)."""
    return (c.p - 1, 2)


def to_jac(pt: Affine) -> Jac:
    if pt is None:
        return JAC_ID
    return (pt[0], pt[1], 1)


def to_affine(c: Curve, pt: Jac) -> Affine:
    X, Y, Z = pt
    if Z % c.p == 0:
        return None
    zi = inv(Z, c.p)
    zi2 = zi * zi % c.p
    return (X * zi2 % c.p, Y * zi2 * zi % c.p)


def jac_double(c: Curve, pt: Jac) -> Jac:
    p = c.p
    X, Y, Z = pt
    if Z == 0 or Y == 0:
        return JAC_ID
    A = X * X % p
    B = Y * Y % p
    C = B * B % p
    D = 2 * ((X + B) * (X + B) - A - C) % p
    E = 3 * A % p
    F = E * E % p
    X3 = (F - 2 * D) % p
    Y3 = (E * (D - X3) - 8 * C) % p
    Z3 = 2 * Y * Z % p
    return (X3, Y3, Z3)


def jac_add(c: Curve, a: Jac, b: Jac) -> Jac:
    p = c.p
    X1, Y1, Z1 = a
    X2, Y2, Z2 = b
    if Z1 == 0:
        return b
    if Z2 == 0:
        return a
    Z1Z1 = Z1 * Z1 % p
    Z2Z2 = Z2 * Z2 % p
    U1 = X1 * Z2Z2 % p
    U2 = X2 * Z1Z1 % p
    S1 = Y1 * Z2 * Z2Z2 % p
    S2 = Y2 * Z1 * Z1Z1 % p
    if U1 == U2:
        if S1 == S2:
            return jac_double(c, a)
        return JAC_ID
    H = (U2 - U1) % p
    R = (S2 - S1) % p
    HH = H * H % p
    HHH = H * HH % p
    V = U1 * HH % p
    X3 = (R * R - HHH - 2 * V) % p
    Y3 = (R * (V - X3) - S1 * HHH) % p
    Z3 = Z1 * Z2 * H % p
    return (X3, Y3, Z3)


def jac_neg(c: Curve, a: Jac) -> Jac:
    return (a[0], (-a[1]) % c.p, a[2])


def jac_eq(c: Curve, a: Jac, b: Jac) -> bool:
    return to_affine(c, a) == to_affine(c, b)


def scalar_mul(c: Curve, k: int, pt: Affine) -> Jac:
    """Plain left-to-right double-and-add -- the 'naive' side of the reference's
    test_multiexp (arithmetic.rs:440-458)."""
    k %= c.r
    acc = JAC_ID
    base = to_jac(pt)
    for bit in bin(k)[2:] if k else "":
        acc = jac_double(c, acc)
        if bit == "1":
            acc = jac_add(c, acc, base)
    return acc


def naive_msm(c: Curve, coeffs: Sequence[int], bases: Sequence[Affine]) -> Jac:
    assert len(coeffs) == len(bases)
    acc = JAC_ID
    for k, b in zip(coeffs, bases):
        acc = jac_add(c, acc, scalar_mul(c, k, b))
    return acc


def batch_normalize(c: Curve, pts: Sequence[Jac]) -> List[Affine]:
    """Montgomery-trick batch inversion (group::Curve::batch_normalize)."""
    p = c.p
    prefix = []
    acc = 1
    for X, Y, Z in pts:
        prefix.append(acc)
        if Z % p:
            acc = acc * Z % p
    acc = inv(acc, p)
    out: List[Affine] = [None] * len(pts)
    for i in range(len(pts) - 1, -1, -1):
        X, Y, Z = pts[i]
        if Z % p == 0:
            continue
        zi = acc * prefix[i] % p
        acc = acc * Z % p
        zi2 = zi * zi % p
        out[i] = (X * zi2 % p, Y * zi2 * zi % p)
    return out


# --------------------------------------------------------------------------------------
# best_multiexp -- arithmetic.rs:143-180
# --------------------------------------------------------------------------------------
def multiexp_window_bits(n: int) -> int:
    """arithmetic.rs:146-152."""
    if n < 4:
        return 1
    if n < 32:
        return 3
    return int(math.ceil(math.log(float(n))))


def get_at(segment: int, c_bits: int, repr32: bytes) -> int:
    """arithmetic.rs:95-111: c bits at bit offset segment*c of the 32-byte LE repr."""
    skip_bits = segment * c_bits
    skip_bytes = skip_bits // 8
    if skip_bytes >= 32:
        return 0
    v = repr32[skip_bytes : skip_bytes + 8].ljust(8, b"\0")
    tmp = int.from_bytes(v, "little")
    tmp >>= skip_bits - skip_bytes * 8
    return tmp % (1 << c_bits)


def _bucket_sum(c: Curve, c_bits: int, reprs: Sequence[bytes], bases: Sequence[Affine], i: int) -> Jac:
    """Buckets::sum, arithmetic.rs:74-93 (None/Affine/Projective promotion collapses to
    Jacobian adds: same group element)."""
    buckets: List[Jac] = [JAC_ID] * ((1 << c_bits) - 1)
    for rep, base in zip(reprs, bases):
        seg = get_at(i, c_bits, rep)
        if seg != 0:
            buckets[seg - 1] = jac_add(c, buckets[seg - 1], to_jac(base))
    acc = JAC_ID
    run = JAC_ID
    for b in reversed(buckets):
        run = jac_add(c, b, run)
        acc = jac_add(c, acc, run)
    return acc


def best_multiexp(c: Curve, coeffs: Sequence[int], bases: Sequence[Affine], num_threads: int = 8) -> Jac:
    """arithmetic.rs:143-180, both branches.  Panics (AssertionError) on length mismatch
    like the reference's assert_eq! at :144."""
    assert len(coeffs) == len(bases)
    n = len(bases)
    c_bits = multiexp_window_bits(n)
    windows = 256 // c_bits + 1
    reprs = [int(k % c.r).to_bytes(32, "little") for k in coeffs]
    if n > num_threads:
        total = JAC_ID
        for i in reversed(range(windows)):
            acc = _bucket_sum(c, c_bits, reprs, bases, i)
            for _ in range(c_bits * i):
                acc = jac_double(c, acc)
            total = jac_add(c, total, acc)
        return total
    total = JAC_ID
    for i in reversed(range(windows)):
        for _ in range(c_bits):
            total = jac_double(c, total)
        total = jac_add(c, total, _bucket_sum(c, c_bits, reprs, bases, i))
    return total


def small_multiexp(c: Curve, coeffs: Sequence[int], bases: Sequence[Affine]) -> Jac:
    """arithmetic.rs:116-136."""
    reprs = [int(k % c.r).to_bytes(32, "little") for k in coeffs]
    acc = JAC_ID
    for byte_idx in reversed(range(32)):
        for bit_idx in reversed(range(8)):
            acc = jac_double(c, acc)
            for rep, base in zip(reprs, bases):
                if (rep[byte_idx] >> bit_idx) & 1:
                    acc = jac_add(c, acc, to_jac(base))
    return acc


# --------------------------------------------------------------------------------------
# best_fft -- arithmetic.rs:192-295.  The butterfly NETWORK, valid for any omega
# (benches/fft.rs:17 passes a random omega), not "the DFT".
# --------------------------------------------------------------------------------------
def bitreverse(n: int, l: int) -> int:
    r = 0
    for _ in range(l):
        r = (r << 1) | (n & 1)
        n >>= 1
    return r


def best_fft(field: str, a: List[int], omega: int, log_n: int) -> None:
    """In place.  Iterative form, arithmetic.rs:207-251."""
    m = FIELDS[field]
    n = len(a)
    assert n == 1 << log_n  # arithmetic.rs:205
    for k in range(n):
        rk = bitreverse(k, log_n)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    twiddles = [1] * (n // 2)
    for i in range(1, n // 2):
        twiddles[i] = twiddles[i - 1] * omega % m
    chunk = 2
    twiddle_chunk = n // 2
    for _ in range(log_n):
        half = chunk // 2
        for base in range(0, n, chunk):
            t = a[base + half]
            a[base + half] = (a[base] - t) % m
            a[base] = (a[base] + t) % m
            for i in range(1, half):
                t = a[base + half + i] * twiddles[i * twiddle_chunk] % m
                u = a[base + i]
                a[base + i] = (u + t) % m
                a[base + half + i] = (u - t) % m
        chunk *= 2
        twiddle_chunk //= 2


def best_fft_recursive(field: str, a: List[int], omega: int, log_n: int) -> None:
    """Recursive form, arithmetic.rs:253,258-295 (same network)."""
    m = FIELDS[field]
    n = len(a)
    assert n == 1 << log_n
    for k in range(n):
        rk = bitreverse(k, log_n)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    twiddles = [1] * max(1, n // 2)
    for i in range(1, n // 2):
        twiddles[i] = twiddles[i - 1] * omega % m

    def rec(lo: int, nn: int, tc: int) -> None:
        if nn == 2:
            t = a[lo + 1]
            a[lo + 1] = (a[lo] - t) % m
            a[lo] = (a[lo] + t) % m
            return
        h = nn // 2
        rec(lo, h, tc * 2)
        rec(lo + h, h, tc * 2)
        t = a[lo + h]
        a[lo + h] = (a[lo] - t) % m
        a[lo] = (a[lo] + t) % m
        for i in range(1, h):
            t = a[lo + h + i] * twiddles[i * tc] % m
            u = a[lo + i]
            a[lo + i] = (u + t) % m
            a[lo + h + i] = (u - t) % m

    if n >= 2:
        rec(0, n, 1)


def ec_fft(c: Curve, a: List[Jac], omega: int, log_n: int) -> None:
    """best_fft with G = curve point (arithmetic.rs:17-27), as used by Params::new
    (poly/commitment.rs:81-82).  '*' is scalar multiplication."""
    n = len(a)
    assert n == 1 << log_n
    r = c.r
    for k in range(n):
        rk = bitreverse(k, log_n)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    twiddles = [1] * max(1, n // 2)
    for i in range(1, n // 2):
        twiddles[i] = twiddles[i - 1] * omega % r
    chunk = 2
    tc = n // 2
    for _ in range(log_n):
        half = chunk // 2
        for base in range(0, n, chunk):
            for i in range(half):
                t = a[base + half + i]
                if i:
                    aff = to_affine(c, t)
                    t = scalar_mul(c, twiddles[i * tc], aff)
                u = a[base + i]
                a[base + i] = jac_add(c, u, t)
                a[base + half + i] = jac_add(c, u, jac_neg(c, t))
        chunk *= 2
        tc //= 2


# --------------------------------------------------------------------------------------
# parallelize chunking -- arithmetic.rs:345-362 (documented; the oracle runs serially)
# --------------------------------------------------------------------------------------
def parallelize_chunks(n: int, threads: int) -> List[Tuple[int, int]]:
    chunk = n // threads
    if chunk < threads:
        chunk = n
    out = []
    start = 0
    while start < n:
        out.append((start, min(chunk, n - start)))
        start += chunk
    return out


# --------------------------------------------------------------------------------------
# EvaluationDomain -- poly/domain.rs:40-146 and the three transforms
# --------------------------------------------------------------------------------------
class EvaluationDomain:
    def __init__(self, field: str, j: int, k: int, zeta: Optional[int] = None):
        m = FIELDS[field]
        self.field = field
        self.m = m
        self.k = k
        self.n = 1 << k
        self.quotient_poly_degree = j - 1
        ext_k = k
        while (1 << ext_k) < self.n * self.quotient_poly_degree:
            ext_k += 1
        assert ext_k <= S_2ADICITY  # domain.rs:56
        self.extended_k = ext_k
        ew = root_of_unity(field)
        for _ in range(ext_k, S_2ADICITY):
            ew = ew * ew % m
        self.extended_omega = ew
        w = ew
        for _ in range(k, ext_k):
            w = w * w % m
        self.omega = w
        self.omega_inv = inv(w, m)
        self.extended_omega_inv = inv(ew, m)
        self.g_coset = zeta_candidates(field)[0] if zeta is None else zeta
        self.g_coset_inv = self.g_coset * self.g_coset % m
        self.ifft_divisor = inv((1 << k) % m, m)
        self.extended_ifft_divisor = inv((1 << ext_k) % m, m)
        # t(X) = X^n - 1 over the coset (domain.rs:88-107), inverted (:121-128)
        orig = pow(self.g_coset, self.n, m)
        step = pow(ew, self.n, m)
        cur = orig
        t = []
        while True:
            t.append(cur)
            cur = cur * step % m
            if cur == orig:
                break
        assert len(t) == 1 << (ext_k - k)
        self.t_evaluations = [inv((x - 1) % m, m) for x in t]

    def extended_len(self) -> int:
        return 1 << self.extended_k

    def rotate_omega(self, value: int, rotation: int) -> int:
        """domain.rs:408-418."""
        return value * pow(self.omega if rotation >= 0 else self.omega_inv, abs(rotation), self.m) % self.m

    def l_i_range(self, x: int, xn: int, rotations: Sequence[int]) -> List[int]:
        """domain.rs:447-472: results[i] = rotate_omega((x - omega^rot)^-1 * (xn - 1) * barycentric_weight, rot)."""
        m = self.m
        results = [(x - self.rotate_omega(1, r)) % m for r in rotations]
        results = [inv(v, m) if v else 0 for v in results]        # batch_invert (:462): zeros stay zero
        common = (xn - 1) * self.ifft_divisor % m                 # barycentric_weight = 1 / n (:118-128)
        return [self.rotate_omega(v * common % m, r) for v, r in zip(results, rotations)]

    def distribute_powers_zeta(self, a: List[int], into_coset: bool) -> None:
        """domain.rs:357-373."""
        cp = [self.g_coset, self.g_coset_inv] if into_coset else [self.g_coset_inv, self.g_coset]
        for idx in range(len(a)):
            i = idx % 3
            if i:
                a[idx] = a[idx] * cp[i - 1] % self.m

    def ifft(self, a: List[int], omega_inv: int, log_n: int, divisor: int) -> None:
        """domain.rs:375-383."""
        best_fft(self.field, a, omega_inv, log_n)
        for i in range(len(a)):
            a[i] = a[i] * divisor % self.m

    def lagrange_to_coeff(self, a: Sequence[int]) -> List[int]:
        """domain.rs:227-237."""
        assert len(a) == 1 << self.k
        a = list(a)
        self.ifft(a, self.omega_inv, self.k, self.ifft_divisor)
        return a

    def coeff_to_extended(self, a: Sequence[int]) -> List[int]:
        """domain.rs:241-255."""
        assert len(a) == 1 << self.k
        a = list(a)
        self.distribute_powers_zeta(a, True)
        a += [0] * (self.extended_len() - len(a))
        best_fft(self.field, a, self.extended_omega, self.extended_k)
        return a

    def extended_to_coeff(self, a: Sequence[int]) -> List[int]:
        """domain.rs:303-325."""
        assert len(a) == self.extended_len()
        a = list(a)
        self.ifft(a, self.extended_omega_inv, self.extended_k, self.extended_ifft_divisor)
        self.distribute_powers_zeta(a, False)
        return a[: self.n * self.quotient_poly_degree]

    def divide_by_vanishing_poly(self, a: Sequence[int]) -> List[int]:
        """domain.rs:329-348."""
        assert len(a) == self.extended_len()
        t = self.t_evaluations
        return [x * t[i % len(t)] % self.m for i, x in enumerate(a)]


def ast_evaluate(d: "EvaluationDomain", basis: str, ast, polys: Sequence[Sequence[int]]) -> List[int]:
    """Evaluator::evaluate (poly/evaluator.rs:129-228), element by element.  `ast` is a nested tuple:
    ("poly", index, rotation) | ("add", a, b) | ("mul", a, b) | ("scale", a, s) | ("dp", [terms], base) | ("lin", s) | ("const", s);
    basis "lagrange" (rotation = 1 position, linear term s w^i, :538-555) or "extended" (rotation = 2^(extended_k - k) positions,
    poly/domain.rs:286-295; linear term s zeta extended_omega^i, :584-604)."""
    m = d.m
    n = d.n if basis == "lagrange" else d.extended_len()
    stride = 1 if basis == "lagrange" else 1 << (d.extended_k - d.k)
    w = d.omega if basis == "lagrange" else d.extended_omega
    lin0 = 1 if basis == "lagrange" else d.g_coset

    def rec(a) -> List[int]:
        k = a[0]
        if k == "poly":
            rot = a[2] * stride                      # rotate_left for rot >= 0, rotate_right for rot < 0 (poly.rs:237-290)
            return [polys[a[1]][(i + rot) % n] for i in range(n)]
        if k == "add":
            x, y = rec(a[1]), rec(a[2])
            return [(p + q) % m for p, q in zip(x, y)]
        if k == "mul":
            x, y = rec(a[1]), rec(a[2])
            return [p * q % m for p, q in zip(x, y)]
        if k == "scale":
            return [p * a[2] % m for p in rec(a[1])]
        if k == "dp":
            acc = [0] * n
            for term in a[1]:
                t = rec(term)
                acc = [(p * a[2] + q) % m for p, q in zip(acc, t)]
            return acc
        if k == "lin":
            out, cur = [], lin0 * a[1] % m
            for _ in range(n):
                out.append(cur)
                cur = cur * w % m
            return out
        if k == "const":
            return [a[1] % m] * n
        raise ValueError(k)

    return rec(ast)


def eval_polynomial(field: str, poly: Sequence[int], point: int) -> int:
    """arithmetic.rs:298-303."""
    m = FIELDS[field]
    acc = 0
    for coeff in reversed(poly):
        acc = (acc * point + coeff) % m
    return acc


# --------------------------------------------------------------------------------------
# Params -- poly/commitment.rs:26-150.  Generators are SYNTHETIC ([s_i]*(-1,2) from the
# seeded PRNG) because hash_to_curve lives in un-vendored pasta_curves; everything
# downstream (EC-FFT for g_lagrange, commit, commit_lagrange) follows the reference.
# --------------------------------------------------------------------------------------
class Params:
    def __init__(self, curve: Curve, k: int, seed: int = 0x48414C4F32):
        assert k < 32  # commitment.rs:41
        self.curve = curve
        self.k = k
        self.n = 1 << k
        rng = Xoshiro256(seed)
        g0 = generator(curve)
        g_proj = [scalar_mul(curve, rng.field_element(curve.r), g0) for _ in range(self.n)]
        self.g = batch_normalize(curve, g_proj)
        # commitment.rs:77-94: alpha_inv = ROOT_OF_UNITY_INV^(2^(S-k)); EC-FFT; * 2^-k
        r = curve.r
        alpha_inv = inv(root_of_unity(curve.scalar), r)
        for _ in range(k, S_2ADICITY):
            alpha_inv = alpha_inv * alpha_inv % r
        gl = list(g_proj)
        ec_fft(curve, gl, alpha_inv, k)
        minv = pow(inv(2, r), k, r)
        gl = [scalar_mul(curve, minv, to_affine(curve, pt)) for pt in gl]
        self.g_lagrange = batch_normalize(curve, gl)
        self.w = to_affine(curve, scalar_mul(curve, rng.field_element(r), g0))
        self.u = to_affine(curve, scalar_mul(curve, rng.field_element(r), g0))

    @classmethod
    def new(cls, curve: Curve, k: int) -> "Params":
        """Params::new(k) proper (poly/commitment.rs:38-114): generators from hash_to_curve("Halo2-Parameters")."""
        g, w, u = params_generators(curve, k)
        return cls.from_generators(curve, k, g, w, u)

    @classmethod
    def from_generators(cls, curve: Curve, k: int, g: Sequence[Affine], w: Affine, u: Affine) -> "Params":
        """commitment.rs:74-101 from given generators: g_lagrange = batch_normalize(2^-k * EC-iFFT(g))."""
        assert k < 32 and len(g) == 1 << k
        self = cls.__new__(cls)
        self.curve, self.k, self.n = curve, k, 1 << k
        self.g = list(g)
        r = curve.r
        alpha_inv = inv(root_of_unity(curve.scalar), r)
        for _ in range(k, S_2ADICITY):
            alpha_inv = alpha_inv * alpha_inv % r
        gl = [to_jac(pt) for pt in g]
        ec_fft(curve, gl, alpha_inv, k)
        minv = pow(inv(2, r), k, r)
        gl = [scalar_mul(curve, minv, to_affine(curve, pt)) for pt in gl]
        self.g_lagrange = batch_normalize(curve, gl)
        self.w, self.u = w, u
        return self

    def commit(self, poly: Sequence[int], blind: int) -> Jac:
        """commitment.rs:119-130."""
        return best_multiexp(self.curve, list(poly) + [blind], list(self.g) + [self.w])

    def commit_lagrange(self, poly: Sequence[int], blind: int) -> Jac:
        """commitment.rs:135-150."""
        return best_multiexp(self.curve, list(poly) + [blind], list(self.g_lagrange) + [self.w])


# --------------------------------------------------------------------------------------
# Seeded PRNG shared by the oracle, the tests and bench.py (splitmix64 -> xoshiro256**),
# SURVEY.md section 8(d).  Seed convention: 0x48414C4F32 ("HALO2") + config index.
# --------------------------------------------------------------------------------------
MASK64 = (1 << 64) - 1


class Xoshiro256:
    def __init__(self, seed: int):
        s = seed & MASK64
        st = []
        for _ in range(4):
            s = (s + 0x9E3779B97F4A7C15) & MASK64
            z = s
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
            st.append(z ^ (z >> 31))
        self.s = st

    @staticmethod
    def _rotl(x: int, k: int) -> int:
        return ((x << k) | (x >> (64 - k))) & MASK64

    def next_u64(self) -> int:
        s = self.s
        result = (self._rotl((s[1] * 5) & MASK64, 7) * 9) & MASK64
        t = (s[1] << 17) & MASK64
        s[2] ^= s[0]
        s[3] ^= s[1]
        s[1] ^= s[2]
        s[0] ^= s[3]
        s[2] ^= t
        s[3] = self._rotl(s[3], 45)
        return result

    def field_element(self, modulus: int) -> int:
        """Uniform canonical element: 255-bit rejection sampling (limb 0 drawn first)."""
        while True:
            limbs = [self.next_u64() for _ in range(4)]
            v = limbs[0] | (limbs[1] << 64) | (limbs[2] << 128) | ((limbs[3] & ((1 << 63) - 1)) << 192)
            if v < modulus:
                return v


def gen_scalars(field: str, seed: int, n: int) -> List[int]:
    rng = Xoshiro256(seed)
    m = FIELDS[field]
    return [rng.field_element(m) for _ in range(n)]


def gen_points(c: Curve, seed: int, n: int) -> List[Affine]:
    """P_0 = [s]G, P_{i+1} = P_i + [t]G with s, t from the PRNG: n distinct
    pseudo-random points for the price of n additions."""
    rng = Xoshiro256(seed)
    g0 = generator(c)
    cur = scalar_mul(c, rng.field_element(c.r), g0)
    step = scalar_mul(c, rng.field_element(c.r), g0)
    pts = []
    for _ in range(n):
        pts.append(cur)
        cur = jac_add(c, cur, step)
    return batch_normalize(c, pts)


# byte helpers (canonical 32-byte LE, identity = 64 zero bytes) -------------------------
def fe_to_bytes(x: int) -> bytes:
    return int(x).to_bytes(32, "little")


def fe_from_bytes(b: bytes) -> int:
    return int.from_bytes(b, "little")


def affine_to_bytes(pt: Affine) -> bytes:
    if pt is None:
        return b"\0" * 64
    return fe_to_bytes(pt[0]) + fe_to_bytes(pt[1])


def affine_from_bytes(b: bytes) -> Affine:
    x = fe_from_bytes(b[:32])
    y = fe_from_bytes(b[32:64])
    if x == 0 and y == 0:
        return None
    return (x, y)


# --------------------------------------------------------------------------------------
# IPA opening proof -- poly/commitment/prover.rs:27-152, with the transcript factored out:
# the Fiat-Shamir challenges (xi, z, u_j) and the prover's randomness are INPUTS, the points
# and scalars the prover would write to the transcript are OUTPUTS.  (The Blake2b transcript
# itself, transcript.rs, is out of scope -- SURVEY.md section 2.)
# --------------------------------------------------------------------------------------
# Point compression (C::to_bytes / C::from_bytes; book/src/background/curves.md:203-240) and the
# Params wire format built on it (poly/commitment.rs:168-205).  The encoding is the x coordinate, 32 bytes
# little-endian, with the LSB of y in the top bit of the last byte; the identity is 32 zero bytes.
# --------------------------------------------------------------------------------------
def fe_sqrt(field: str, a: int) -> Optional[int]:
    """A square root of a in the field, or None.  Tonelli-Shanks over the 2^32-order subgroup generated by ROOT_OF_UNITY
    (book/src/background/fields.md: both fields have 2-adicity 32); which of the two roots comes back is irrelevant to the
    callers, who fix the sign afterwards."""
    m = FIELDS[field]
    a %= m
    if a == 0:
        return 0
    t = (m - 1) >> S_2ADICITY
    w = pow(a, (t - 1) // 2, m)
    x = a * w % m            # a^((t+1)/2)
    b = x * w % m            # a^t
    z = root_of_unity(field)
    v = S_2ADICITY
    while b != 1:
        k, b2 = 0, b
        while b2 != 1:
            b2 = b2 * b2 % m
            k += 1
            if k == v:
                return None  # a is not a square
        zz = z
        for _ in range(v - k - 1):
            zz = zz * zz % m
        x = x * zz % m
        z = zz * zz % m
        b = b * z % m
        v = k
    assert x * x % m == a
    return x


def compress(pt: Affine) -> bytes:
    """C::to_bytes (curves.md:203-225)."""
    if pt is None:
        return b"\0" * 32
    return (pt[0] | ((pt[1] & 1) << 255)).to_bytes(32, "little")


def decompress(c: Curve, b: bytes) -> Affine:
    """C::from_bytes (curves.md:227-240).  Raises ValueError on an invalid encoding (C::read returns io::Error)."""
    v = int.from_bytes(b, "little")
    ysign, x = v >> 255, v & ((1 << 255) - 1)
    if x == 0:
        if ysign:
            raise ValueError("invalid point encoding: x = 0 with the sign bit set")
        return None
    if x >= c.p:
        raise ValueError("invalid point encoding: x is not a canonical field element")
    y = fe_sqrt(c.base, (x * x % c.p * x + 5) % c.p)
    if y is None:
        raise ValueError("invalid point encoding: x^3 + 5 is not a square")
    if (y & 1) != ysign:
        y = c.p - y
    return (x, y)


def params_to_bytes(k: int, g: Sequence[Affine], g_lagrange: Sequence[Affine], w: Affine, u: Affine) -> bytes:
    """Params::write (poly/commitment.rs:168-181)."""
    return k.to_bytes(4, "little") + b"".join(compress(p) for p in list(g) + list(g_lagrange) + [w, u])


def params_from_bytes(c: Curve, data: bytes):
    """Params::read (poly/commitment.rs:183-205) -> (k, g, g_lagrange, w, u)."""
    k = int.from_bytes(data[:4], "little")
    n = 1 << k
    if len(data) < 4 + 32 * (2 * n + 2):
        raise ValueError("unexpected end of file")       # read_exact
    pts = [decompress(c, data[4 + 32 * i:4 + 32 * (i + 1)]) for i in range(2 * n + 2)]
    return k, pts[:n], pts[n:2 * n], pts[2 * n], pts[2 * n + 1]


# --------------------------------------------------------------------------------------
def kate_division(field: str, a: Sequence[int], b: int) -> List[int]:
    """arithmetic.rs:322-341: the quotient of a(X) by (X - b), remainder dropped; len(a) - 1 coefficients."""
    m = FIELDS[field]
    nb = (-b) % m
    q = [0] * (len(a) - 1)
    tmp = 0
    for qi, r in zip(range(len(q) - 1, -1, -1), reversed(list(a))):
        lead = (r - tmp) % m
        q[qi] = lead
        tmp = lead * nb % m
    return q


def compute_inner_product(m: int, a: Sequence[int], b: Sequence[int]) -> int:
    """arithmetic.rs:308-319."""
    assert len(a) == len(b)
    return sum(x * y for x, y in zip(a, b)) % m


def parallel_generator_collapse(c: Curve, g: List[Affine], challenge: int) -> List[Affine]:
    """poly/commitment/prover.rs:154-166: g_lo[i] + [challenge] g_hi[i], batch-normalised."""
    half = len(g) // 2
    tmp = [jac_add(c, to_jac(g[i]), scalar_mul(c, challenge, g[i + half])) for i in range(half)]
    return batch_normalize(c, tmp)


def ipa_rounds(c: Curve, g: Sequence[Affine], w: Affine, u: Affine, p_prime: Sequence[int], x3: int, z: int,
               challenges: Sequence[int], l_rand: Sequence[int], r_rand: Sequence[int]):
    """The round loop of commitment::create_proof (prover.rs:100-142) for a given p_prime (prover.rs:80),
    evaluation point x3, challenge z and per-round challenges u_j / randomness.  Returns
    ([L_j affine], [R_j affine], c = final p_prime[0], sum_j (l_rand_j / u_j + r_rand_j u_j))."""
    r = c.r
    n = len(g)
    k = n.bit_length() - 1
    assert n == 1 << k and len(p_prime) == n and len(challenges) == k
    p_prime = [x % r for x in p_prime]
    b = [1] * n
    for i in range(1, n):
        b[i] = b[i - 1] * x3 % r
    g_prime = list(g)
    ls, rs = [], []
    f_delta = 0
    for j in range(k):
        half = 1 << (k - j - 1)
        l_j = best_multiexp(c, p_prime[half:], g_prime[:half])
        r_j = best_multiexp(c, p_prime[:half], g_prime[half:])
        value_l = compute_inner_product(r, p_prime[half:], b[:half])
        value_r = compute_inner_product(r, p_prime[:half], b[half:])
        l_j = jac_add(c, l_j, best_multiexp(c, [value_l * z % r, l_rand[j]], [u, w]))
        r_j = jac_add(c, r_j, best_multiexp(c, [value_r * z % r, r_rand[j]], [u, w]))
        ls.append(to_affine(c, l_j))
        rs.append(to_affine(c, r_j))
        u_j = challenges[j] % r
        u_inv = inv(u_j, r)
        for i in range(half):
            p_prime[i] = (p_prime[i] + p_prime[i + half] * u_inv) % r
            b[i] = (b[i] + b[i + half] * u_j) % r
        p_prime = p_prime[:half]
        b = b[:half]
        g_prime = parallel_generator_collapse(c, g_prime, u_j)
        f_delta = (f_delta + l_rand[j] * u_inv + r_rand[j] * u_j) % r
    assert len(p_prime) == 1
    return ls, rs, p_prime[0], f_delta


# --------------------------------------------------------------------------------------
# hash_to_curve -- C::CurveExt::hash_to_curve(domain_prefix)(message), the call at
# poly/commitment.rs:52,102.  The implementation lives in the un-vendored crate
# pasta_curves 0.5.1 (src/hashtocurve.rs, src/curves.rs); it is the hash-to-curve suite
# "<curve>_XMD:BLAKE2b_SSWU_RO_" of the IETF hash-to-curve specification (RFC 9380):
#   hash_to_field   = expand_message_xmd (section 5.3.1) with BLAKE2b-512 (block 128 B, all-zero
#                     personalisation), DST = domain_prefix || "-" || curve_id || "_XMD:BLAKE2b_SSWU_RO_",
#                     two 64-byte chunks, each read big-endian and reduced mod the base field;
#   map_to_curve    = simplified SWU (section 6.6.2) onto the curve iso-<curve>: y^2 = x^3 + A x + 1265
#                     that is 3-isogenous to y^2 = x^3 + 5, with Z = -13 and sgn0 = parity;
#   the two images are ADDED on the iso curve, then mapped through the 3-isogeny (section 6.6.3).
# Constants are not copied from anywhere: the iso curve is the codomain of Velu's 3-isogeny from
# y^2 = x^3 + 5 with kernel x0, x0^3 = -20 (giving A = -30 x0^2, B = 5 + 7*180 = 1265), and the isogeny
# back is its dual (x-leading coefficient 1/9), both derived below.  Which of the three cube roots
# pasta uses is fixed by `ISO_A`; the whole construction is PINNED by the reference's golden
# commitments (tests/plonk_api.rs:958-982: fixed_commitments[2] is an all-zero column committed with
# Blind::default() = 1, i.e. the point w = hasher(&[1]); the other entries are MSMs over g_lagrange,
# i.e. over all 32 hashed generators) -- tests/test_oracle_golden.py.
# --------------------------------------------------------------------------------------
ISO_A = {
    "pallas": 0x18354A2EB0EA8C9C49BE2D7258370742B74134581A27A59F92BB4B0B657A014B,
    "vesta": 0x267F9B2EE592271A81639C4D96F787739673928C7D01B212C515AD7242EAA6B1,
}
ISO_B = 1265
SWU_Z = -13


def _cube_roots(a: int, m: int) -> List[int]:
    """All cube roots of a modulo a prime m = 1 (mod 3), m - 1 = 3^s t."""
    a %= m
    if pow(a, (m - 1) // 3, m) != 1:
        return []
    s, t = 0, m - 1
    while t % 3 == 0:
        s, t = s + 1, t // 3
    e, mult = ((t + 1) // 3, 1) if t % 3 == 2 else ((2 * t + 1) // 3, 2)
    r = pow(a, e, m)              # r^3 = a * a^(mult t)
    b = pow(a, mult * t, m)       # in the 3-Sylow subgroup, a cube there
    z = pow(MULT_GEN, t, m)       # generates the 3-Sylow subgroup (order 3^s)
    z3, zz, j = z * z * z % m, 1, 0
    while zz != b:                # b = z^(3j); s is tiny for both Pasta fields
        zz, j = zz * z3 % m, j + 1
        assert j < 3 ** s
    r = r * pow(z, 3 ** s - j, m) % m
    assert pow(r, 3, m) == a
    g = pow(MULT_GEN, (m - 1) // 3, m)
    return [r, r * g % m, r * g * g % m]


_ISO_CACHE: dict = {}


def iso_constants(c: Curve) -> dict:
    """Derive iso-<curve> and the 3-isogeny iso-<curve> -> <curve> from first principles.

    E: y^2 = x^3 + 5.  psi_3(E) = 3x(x^3 + 20): the kernels of the 3-isogenies with j != 0 codomain are
    {O, (x0, +-y0)} with x0^3 = -20 (y0^2 = -15; only y0^2 enters).  Velu: t = 6 x0^2, u = 4 y0^2 = -60,
    w = u + x0 t = -180, codomain y^2 = x^3 - 5t x + (5 - 7w) = x^3 - 30 x0^2 x + 1265.
    The dual isogeny has kernel phi(E[3]) = {O, (xk, .)} with xk = phi_x(0) = -t/x0 + u/x0^2; Velu from the iso
    curve with that kernel lands on y^2 = x^3 + 3^6 * 5, and (x/9, y/27) brings it to E:
        x' = (x + T/(x - xk) + U/(x - xk)^2) / 9,     y' = y (1 - T/(x - xk)^2 - 2U/(x - xk)^3) / 27.
    """
    if c.name in _ISO_CACHE:
        return _ISO_CACHE[c.name]
    m = c.p
    A = ISO_A[c.name]
    x0 = [x for x in _cube_roots(-20, m) if (-30 * x * x) % m == A]
    assert len(x0) == 1, "ISO_A is not a Velu codomain of y^2 = x^3 + 5"
    x0 = x0[0]
    t, u = 6 * x0 * x0 % m, (-60) % m
    xk = (-t * inv(x0, m) + u * inv(x0 * x0 % m, m)) % m
    T = (6 * xk * xk + 2 * A) % m
    U = 4 * (xk * xk * xk + A * xk + ISO_B) % m
    W = (U + xk * T) % m
    assert (A - 5 * T) % m == 0 and (ISO_B - 7 * W) % m == 729 * CURVE_B % m, "dual isogeny does not land on E"
    k = {"A": A, "B": ISO_B, "xk": xk, "T": T, "U": U, "Z": SWU_Z % m}
    _ISO_CACHE[c.name] = k
    return k


def hash_to_field(c: Curve, domain_prefix: str, message: bytes) -> Tuple[int, int]:
    """RFC 9380 section 5.2/5.3.1 for m = 1, count = 2, L = 64, H = BLAKE2b-512 (r_in_bytes = 128)."""
    import hashlib
    dst = domain_prefix.encode() + b"-" + c.name.encode() + b"_XMD:BLAKE2b_SSWU_RO_"
    assert len(dst) < 256
    dst_prime = dst + bytes([len(dst)])
    h = lambda data: hashlib.blake2b(data, digest_size=64).digest()
    b0 = h(bytes(128) + message + bytes([0, 128, 0]) + dst_prime)
    b1 = h(b0 + b"\x01" + dst_prime)
    b2 = h(bytes(x ^ y for x, y in zip(b0, b1)) + b"\x02" + dst_prime)
    return int.from_bytes(b1, "big") % c.p, int.from_bytes(b2, "big") % c.p


def map_to_curve_simple_swu(c: Curve, u: int) -> Optional[Tuple[int, int]]:
    """RFC 9380 section 6.6.2 onto iso-<curve> (affine result; never the identity)."""
    m = c.p
    k = iso_constants(c)
    A, B, Z = k["A"], k["B"], k["Z"]
    zu2 = Z * u * u % m
    ta = (zu2 * zu2 + zu2) % m
    tv1 = inv(ta, m) if ta else 0                       # inv0
    x1 = (-B * inv(A, m)) % m * (1 + tv1) % m
    if tv1 == 0:
        x1 = B * inv(Z * A % m, m) % m
    gx1 = (x1 * x1 * x1 + A * x1 + B) % m
    y1 = fe_sqrt(c.base, gx1)
    if y1 is not None:
        x, y = x1, y1
    else:
        x = zu2 * x1 % m
        y = fe_sqrt(c.base, (x * x * x + A * x + B) % m)
        assert y is not None
    if (u & 1) != (y & 1):                              # sgn0(u) == sgn0(y)
        y = (m - y) % m
    return x, y


def _iso_add(c: Curve, a: Optional[Tuple[int, int]], b: Optional[Tuple[int, int]]) -> Optional[Tuple[int, int]]:
    """Affine addition on iso-<curve> (y^2 = x^3 + A x + B, A != 0)."""
    m = c.p
    if a is None:
        return b
    if b is None:
        return a
    (x1, y1), (x2, y2) = a, b
    if x1 == x2:
        if (y1 + y2) % m == 0:
            return None
        lam = (3 * x1 * x1 + iso_constants(c)["A"]) * inv(2 * y1 % m, m) % m
    else:
        lam = (y2 - y1) * inv((x2 - x1) % m, m) % m
    x3 = (lam * lam - x1 - x2) % m
    return x3, (lam * (x1 - x3) - y1) % m


def iso_map(c: Curve, pt: Optional[Tuple[int, int]]) -> Affine:
    """The 3-isogeny iso-<curve> -> <curve> (RFC 9380 section 6.6.3; kernel points go to the identity)."""
    if pt is None:
        return None
    m = c.p
    k = iso_constants(c)
    x, y = pt
    d = (x - k["xk"]) % m
    if d == 0:
        return None
    di = inv(d, m)
    di2 = di * di % m
    xo = (x + k["T"] * di + k["U"] * di2) % m * inv(9, m) % m
    yo = y * ((1 - k["T"] * di2 - 2 * k["U"] * di2 % m * di) % m) % m * inv(27, m) % m
    return xo, yo


def hash_to_curve(c: Curve, domain_prefix: str):
    """C::CurveExt::hash_to_curve(domain_prefix) -> closure over the message (poly/commitment.rs:52,102)."""
    def hasher(message: bytes) -> Affine:
        u0, u1 = hash_to_field(c, domain_prefix, message)
        r = _iso_add(c, map_to_curve_simple_swu(c, u0), map_to_curve_simple_swu(c, u1))
        out = iso_map(c, r)
        assert on_curve(c, out)
        return out
    return hasher


def params_generators(c: Curve, k: int) -> Tuple[List[Affine], Affine, Affine]:
    """(g, w, u) of Params::new(k): poly/commitment.rs:46-58 (message = 0 || i as u32 LE) and :102-105."""
    hasher = hash_to_curve(c, "Halo2-Parameters")
    g = [hasher(b"\0" + i.to_bytes(4, "little")) for i in range(1 << k)]
    return g, hasher(b"\x01"), hasher(b"\x02")


def permute_expression_pair(field: str, input_expression: Sequence[int], table_expression: Sequence[int], usable_rows: int):
    """plonk/lookup/prover.rs:563-647 without the blinding rows (:625-627, random): (A', S') over the usable rows, or None
    where the reference returns Error::ConstraintSystemFailure (:605-608).  Line for line: sort the input (:577-581), count the
    table values (:584-591), walk the sorted input -- first row of a run takes its own value and removes one instance from the
    map, repeated rows are remembered (:595-614) -- then hand the leftover table values, ascending, to the remembered rows
    popped from the back (:617-622)."""
    u = int(usable_rows)
    permuted_input = sorted(int(x) % FIELDS[field] for x in input_expression[:u])
    leftover: Dict[int, int] = {}
    for t in table_expression[:u]:
        leftover[int(t)] = leftover.get(int(t), 0) + 1
    permuted_table = [0] * u
    repeated_rows: List[int] = []
    for row, v in enumerate(permuted_input):
        if row == 0 or v != permuted_input[row - 1]:
            permuted_table[row] = v
            if leftover.get(v, 0) > 0:
                leftover[v] -= 1
            else:
                return None
        else:
            repeated_rows.append(row)
    for coeff in sorted(leftover):                     # BTreeMap iteration order: ascending keys
        for _ in range(leftover[coeff]):
            permuted_table[repeated_rows.pop()] = coeff
    assert not repeated_rows
    return permuted_input, permuted_table


# --------------------------------------------------------------------------------------
# The verifier's side of the polynomial commitment scheme: MSM<C> (poly/commitment/msm.rs:9-178), the opening
# verifier commitment::verify_proof + Guard + compute_b + compute_s (poly/commitment/verifier.rs:13-171), and the
# opening prover commitment::create_proof whole (poly/commitment/prover.rs:36-151; ipa_rounds above is its round loop).
# The transcript is an argument (any object with write_point / write_scalar / read_point / read_scalar /
# squeeze_challenge on ints and affine tuples); points are Affine tuples, None = the identity.
# --------------------------------------------------------------------------------------
class VerifyError(Exception):
    """Error::OpeningError / Error::SamplingError of poly/commitment/verifier.rs:77-134."""


def compute_b(m: int, x: int, u: Sequence[int]) -> int:
    """verifier.rs:144-153: prod_{i<k} (1 + u_{k-1-i} x^(2^i))."""
    tmp, cur = 1, x % m
    for u_j in reversed(list(u)):
        tmp = tmp * (1 + u_j * cur) % m
        cur = cur * cur % m
    return tmp


def compute_s(m: int, u: Sequence[int], init: int) -> List[int]:
    """verifier.rs:156-171: the coefficients of prod_{i<k} (1 + u_{k-1-i} X^(2^i)) times init, by the reference's
    doubling copies."""
    u = list(u)
    assert len(u) > 0
    v = [0] * (1 << len(u))
    v[0] = init % m
    for i, u_j in enumerate(reversed(u)):
        ln = 1 << i
        for t in range(ln):
            v[ln + t] = v[t] * u_j % m
    return v


class MSM:
    """poly/commitment/msm.rs:9-178.  `other` maps x -> (scalar, y) like the reference's BTreeMap<C::Base, _>; its
    iteration order (ascending x) only fixes the order of the terms of one multiexp, not the result."""

    def __init__(self, c: Curve, g: Sequence[Affine], w: Affine, u: Affine):
        self.c, self.g, self.w, self.u = c, g, w, u
        self.n = len(g)
        self.g_scalars: Optional[List[int]] = None
        self.w_scalar: Optional[int] = None
        self.u_scalar: Optional[int] = None
        self.other: Dict[int, Tuple[int, int]] = {}

    def clone(self) -> "MSM":
        o = MSM(self.c, self.g, self.w, self.u)
        o.g_scalars = None if self.g_scalars is None else list(self.g_scalars)
        o.w_scalar, o.u_scalar, o.other = self.w_scalar, self.u_scalar, dict(self.other)
        return o

    def _merge(self, x: int, y: int, scalar: int) -> None:     # :40-50, :73-83
        r = self.c.r
        if x in self.other:
            ours, our_y = self.other[x]
            if our_y == y:
                self.other[x] = ((ours + scalar) % r, our_y)
            else:
                assert our_y == (-y) % self.c.p
                self.other[x] = ((ours - scalar) % r, our_y)
        else:
            self.other[x] = (scalar % r, y)

    def add_msm(self, other: "MSM") -> None:                   # :37-62
        for x, (scalar, y) in other.other.items():
            self._merge(x, y, scalar)
        if other.g_scalars is not None:
            self.add_to_g_scalars(other.g_scalars)
        if other.w_scalar is not None:
            self.add_to_w_scalar(other.w_scalar)
        if other.u_scalar is not None:
            self.add_to_u_scalar(other.u_scalar)

    def append_term(self, scalar: int, point: Affine) -> None:  # :65-84 (the identity is skipped)
        if point is not None:
            self._merge(point[0], point[1], scalar)

    def add_constant_term(self, constant: int) -> None:        # :87-95
        if self.g_scalars is None:
            self.g_scalars = [0] * self.n
        self.g_scalars[0] = (self.g_scalars[0] + constant) % self.c.r

    def add_to_g_scalars(self, scalars: Sequence[int]) -> None:  # :99-109
        assert len(scalars) == self.n
        if self.g_scalars is None:
            self.g_scalars = [s % self.c.r for s in scalars]
        else:
            self.g_scalars = [(a + b) % self.c.r for a, b in zip(self.g_scalars, scalars)]

    def add_to_w_scalar(self, scalar: int) -> None:            # :112-114
        self.w_scalar = scalar % self.c.r if self.w_scalar is None else (self.w_scalar + scalar) % self.c.r

    def add_to_u_scalar(self, scalar: int) -> None:            # :117-119
        self.u_scalar = scalar % self.c.r if self.u_scalar is None else (self.u_scalar + scalar) % self.c.r

    def scale(self, factor: int) -> None:                      # :122-135
        r = self.c.r
        if self.g_scalars is not None:
            self.g_scalars = [a * factor % r for a in self.g_scalars]
        self.other = {x: (s * factor % r, y) for x, (s, y) in self.other.items()}
        if self.w_scalar is not None:
            self.w_scalar = self.w_scalar * factor % r
        if self.u_scalar is not None:
            self.u_scalar = self.u_scalar * factor % r

    def terms(self) -> Tuple[List[int], List[Affine]]:
        """The (scalars, bases) of eval's one multiexp in the reference's order (:142-172): other, w, u, g."""
        scalars: List[int] = []
        bases: List[Affine] = []
        for x in sorted(self.other):
            s, y = self.other[x]
            scalars.append(s)
            bases.append((x, y))
        if self.w_scalar is not None:
            scalars.append(self.w_scalar)
            bases.append(self.w)
        if self.u_scalar is not None:
            scalars.append(self.u_scalar)
            bases.append(self.u)
        if self.g_scalars is not None:
            scalars.extend(self.g_scalars)
            bases.extend(self.g)
        return scalars, bases

    def eval(self, multiexp=None) -> bool:                     # :138-177: best_multiexp(...).is_identity()
        scalars, bases = self.terms()
        res = (multiexp or (lambda s, b: best_multiexp(self.c, s, b)))(scalars, bases)
        return res[2] == 0


class Guard:
    """verifier.rs:13-63."""

    def __init__(self, msm: MSM, neg_c: int, u: List[int]):
        self.msm, self.neg_c, self.u = msm, neg_c, u

    def use_challenges(self) -> MSM:                           # :36-41
        self.msm.add_to_g_scalars(compute_s(self.msm.c.r, self.u, self.neg_c))
        return self.msm

    def use_g(self, g: Affine) -> Tuple[MSM, Tuple[Affine, List[int]]]:   # :45-55
        self.msm.append_term(self.neg_c, g)
        return self.msm, (g, list(self.u))

    def compute_g(self, multiexp=None) -> Affine:              # :58-62
        c = self.msm.c
        s = compute_s(c.r, self.u, 1)
        return to_affine(c, (multiexp or (lambda a, b: best_multiexp(c, a, b)))(s, list(self.msm.g)))


def ipa_verify_proof(k: int, msm: MSM, transcript, x: int, v: int) -> Guard:
    """commitment::verify_proof (verifier.rs:67-141): `msm` evaluates to the commitment P being opened at x to v."""
    r = msm.c.r
    msm.add_constant_term((-v) % r)                            # :76  P' = P - [v] G_0 + [xi] S
    try:
        s_poly_commitment = transcript.read_point()            # :77
    except Exception as e:                                     # Error::OpeningError
        raise VerifyError("OpeningError") from e
    xi = transcript.squeeze_challenge()                        # :78
    msm.append_term(xi, s_poly_commitment)                     # :79
    z = transcript.squeeze_challenge()                         # :81
    rounds = []
    for _ in range(k):                                         # :84-93
        try:
            l = transcript.read_point()
            rr = transcript.read_point()
        except Exception as e:
            raise VerifyError("OpeningError") from e
        rounds.append((l, rr, transcript.squeeze_challenge()))
    u: List[int] = []
    for l, rr, u_j in rounds:                                  # :95-111 (batch_invert: the same inverses)
        msm.append_term(inv(u_j, r), l)
        msm.append_term(u_j, rr)
        u.append(u_j)
    try:
        c_val = transcript.read_scalar()                       # :126  Error::SamplingError
        f = transcript.read_scalar()                           # :128
    except Exception as e:
        raise VerifyError("SamplingError") from e
    neg_c = (-c_val) % r
    b = compute_b(r, x, u)                                     # :129
    msm.add_to_u_scalar(neg_c * b % r * z % r)                 # :131
    msm.add_to_w_scalar((-f) % r)                              # :132
    return Guard(msm, neg_c, u)


def ipa_create_proof(c: Curve, g: Sequence[Affine], w: Affine, u: Affine, transcript, p_poly: Sequence[int], p_blind: int, x3: int,
                     s_poly: Sequence[int], s_poly_blind: int, l_rand: Sequence[int], r_rand: Sequence[int]) -> None:
    """commitment::create_proof whole (prover.rs:36-151).  The reference draws s_poly, its blind and the per-round
    randomness from its RNG; here they are arguments (s_poly any polynomial of the same length: its evaluation at x3 is
    removed as at :50-51)."""
    r = c.r
    n = len(g)
    k = n.bit_length() - 1
    assert len(p_poly) == n                                    # :41
    s = [a % r for a in s_poly]
    s[0] = (s[0] - eval_polynomial_mod(r, s, x3)) % r          # :51-52
    transcript.write_point(to_affine(c, best_multiexp(c, s + [s_poly_blind], list(g) + [w])))   # :57-58
    xi = transcript.squeeze_challenge()                        # :63
    z = transcript.squeeze_challenge()                         # :67
    p_prime = [(a * xi + b) % r for a, b in zip(s, p_poly)]    # :71  p' = s * xi + p
    v = eval_polynomial_mod(r, p_prime, x3)                    # :72
    p_prime[0] = (p_prime[0] - v) % r                          # :73
    f = (s_poly_blind * xi + p_blind) % r                      # :74-76
    state = {"f": f}
    challenges: List[int] = []
    # the round loop (:100-142) needs each challenge right after its L_j, R_j: run it round by round
    b = [1] * n
    for i in range(1, n):
        b[i] = b[i - 1] * x3 % r                               # :86-93
    g_prime = list(g)
    for j in range(k):
        half = 1 << (k - j - 1)
        l_j = best_multiexp(c, p_prime[half:], g_prime[:half])
        r_j = best_multiexp(c, p_prime[:half], g_prime[half:])
        value_l = compute_inner_product(r, p_prime[half:], b[:half])
        value_r = compute_inner_product(r, p_prime[:half], b[half:])
        l_j = jac_add(c, l_j, best_multiexp(c, [value_l * z % r, l_rand[j]], [u, w]))
        r_j = jac_add(c, r_j, best_multiexp(c, [value_r * z % r, r_rand[j]], [u, w]))
        transcript.write_point(to_affine(c, l_j))              # :119-120
        transcript.write_point(to_affine(c, r_j))
        u_j = transcript.squeeze_challenge()                   # :122
        u_inv = inv(u_j, r)
        challenges.append(u_j)
        for i in range(half):                                  # :128-133
            p_prime[i] = (p_prime[i] + p_prime[i + half] * u_inv) % r
            b[i] = (b[i] + b[i + half] * u_j) % r
        p_prime, b = p_prime[:half], b[:half]
        g_prime = parallel_generator_collapse(c, g_prime, u_j)  # :136-137
        state["f"] = (state["f"] + l_rand[j] * u_inv + r_rand[j] * u_j) % r   # :140-141
    transcript.write_scalar(p_prime[0])                        # :145-149
    transcript.write_scalar(state["f"])


def eval_polynomial_mod(m: int, poly: Sequence[int], point: int) -> int:
    """arithmetic.rs:297-303 with the modulus given directly."""
    acc = 0
    for coeff in reversed(list(poly)):
        acc = (acc * point + coeff) % m
    return acc


# --------------------------------------------------------------------------------------
# The multi-point opening argument: poly/multiopen.rs:144-275 (construct_intermediate_sets), poly/multiopen/prover.rs:18-124
# (create_proof), poly/multiopen/verifier.rs:14-140 (verify_proof), arithmetic.rs:376-432 (lagrange_interpolate).
# Queries are (point, object, ...) records; "the same polynomial / commitment" is object identity, as the reference's
# PolynomialPointer / CommitmentReference compare by pointer (prover.rs:133-137, multiopen.rs:103-114).  The reference draws
# f's blind and the opening's randomness from its RNG; here `rng` is any object with scalar() -> int and poly(n) -> [int].
# --------------------------------------------------------------------------------------
class ProverQuery:                                             # multiopen.rs:42-50
    def __init__(self, point: int, poly: Sequence[int], blind: int):
        self.point, self.poly, self.blind = point, poly, blind

    def key(self):
        return id(self.poly)

    def value(self):                                           # Query::get_eval: the polynomial itself, with its blind (prover.rs:148-152)
        return (self.poly, self.blind)


class VerifierQuery:                                           # multiopen.rs:53-88
    def __init__(self, commitment, point: int, eval_: int):
        self.commitment, self.point, self.eval = commitment, point, eval_

    @classmethod
    def new_commitment(cls, commitment, point: int, eval_: int) -> "VerifierQuery":
        return cls(commitment, point, eval_)

    new_msm = new_commitment                                   # the commitment is an MSM object instead of a point

    def key(self):
        return id(self.commitment)

    def value(self):
        return self.eval


def construct_intermediate_sets(queries, prover: bool):
    """multiopen.rs:144-275.  Returns (commitment_data, point_sets) with commitment_data a list of dicts {commitment, set_index,
    point_indices, evals} in first-seen order, or None when one commitment is queried twice at one point with different
    evaluations (:243-249; a prover's "evaluation" is the polynomial itself, so its repeats are merely redundant -- and rejected
    alike, since the slot is already filled)."""
    queries = list(queries)
    commitment_map: Dict[int, dict] = {}                       # IndexMap: insertion order
    point_index_map: Dict[int, int] = {}
    for q in queries:                                          # :162-173
        idx = point_index_map.setdefault(q.point, len(point_index_map))
        commitment_map.setdefault(q.key(), {"commitment": q.poly if prover else q.commitment, "blind": getattr(q, "blind", None),
                                            "set_index": 0, "point_indices": [], "evals": []})["point_indices"].append(idx)
    inverse = {i: p for p, i in point_index_map.items()}       # :176-179
    point_idx_sets: Dict[Tuple[int, ...], int] = {}
    commitment_set: Dict[int, Tuple[int, ...]] = {}
    for key, data in commitment_map.items():                   # :186-203
        pset = tuple(sorted(set(data["point_indices"])))
        commitment_set[key] = pset
        point_idx_sets.setdefault(pset, len(point_idx_sets))
        data["evals"] = [None] * len(pset)
    for q in queries:                                          # :206-250
        data = commitment_map[q.key()]
        pset = commitment_set[q.key()]
        data["set_index"] = point_idx_sets[pset]
        slot = pset.index(point_index_map[q.point])
        if data["evals"][slot] is None:
            data["evals"][slot] = q.value()
        else:
            return None
    point_sets: List[List[int]] = [[] for _ in point_idx_sets]  # :266-272
    for pset, sidx in point_idx_sets.items():
        point_sets[sidx] = [inverse[i] for i in pset]
    return list(commitment_map.values()), point_sets


def lagrange_interpolate(m: int, points: Sequence[int], evals: Sequence[int]) -> List[int]:
    """arithmetic.rs:376-432: the coefficients of the polynomial of degree < len(points) through (points[i], evals[i])."""
    assert len(points) == len(evals)
    if len(points) == 1:
        return [evals[0] % m]
    final = [0] * len(points)
    for j, (x_j, ev) in enumerate(zip(points, evals)):
        tmp = [1]
        for kk, x_k in enumerate(points):
            if kk == j:
                continue
            denom = inv((x_j - x_k) % m, m)
            a = tmp + [0]
            b = [0] + tmp
            tmp = [(ai * ((-denom * x_k) % m) + bi * denom) % m for ai, bi in zip(a, b)]
        for i, coeff in enumerate(tmp):
            final[i] = (final[i] + coeff * ev) % m
    return final


def multiopen_create_proof(c: Curve, g: Sequence[Affine], w: Affine, u: Affine, rng, transcript, queries) -> None:
    """poly/multiopen/prover.rs:18-124."""
    r = c.r
    n = len(g)
    x1 = transcript.squeeze_challenge()                        # :38
    x2 = transcript.squeeze_challenge()                        # :39
    sets = construct_intermediate_sets(queries, prover=True)   # :41-46
    if sets is None:
        raise ValueError("queries iterator contains mismatching evaluations")
    poly_map, point_sets = sets
    q_polys: List[Optional[List[int]]] = [None] * len(point_sets)
    q_blinds = [0] * len(point_sets)
    for data in poly_map:                                      # :53-73: q_i = q_i * x_1 + poly, the blinds alike
        s = data["set_index"]
        poly = [a % r for a in data["commitment"]]
        q_polys[s] = poly if q_polys[s] is None else [(a * x1 + b) % r for a, b in zip(q_polys[s], poly)]
        q_blinds[s] = (q_blinds[s] * x1 + data["blind"]) % r
    q_prime: Optional[List[int]] = None
    for points, poly in zip(point_sets, q_polys):              # :75-96
        cur = list(poly)
        for pt in points:
            cur = kate_division_mod(r, cur, pt)
        cur = cur + [0] * (n - len(cur))                       # :84 resize
        q_prime = cur if q_prime is None else [(a * x2 + b) % r for a, b in zip(q_prime, cur)]
    q_prime_blind = rng.scalar()                               # :98
    transcript.write_point(to_affine(c, best_multiexp(c, q_prime + [q_prime_blind], list(g) + [w])))   # :99-101
    x3 = transcript.squeeze_challenge()                        # :103
    for q in q_polys:                                          # :107-109
        transcript.write_scalar(eval_polynomial_mod(r, q, x3))
    x4 = transcript.squeeze_challenge()                        # :111
    p_poly, p_blind = q_prime, q_prime_blind
    for poly, blind in zip(q_polys, q_blinds):                 # :113-121
        p_poly = [(a * x4 + b) % r for a, b in zip(p_poly, poly)]
        p_blind = (p_blind * x4 + blind) % r
    s_poly = rng.poly(n)                                       # commitment::create_proof draws these (prover.rs:46-54, :112-113)
    s_blind = rng.scalar()
    k = n.bit_length() - 1
    rand = [(rng.scalar(), rng.scalar()) for _ in range(k)]
    ipa_create_proof(c, g, w, u, transcript, p_poly, p_blind, x3, s_poly, s_blind, [a for a, _ in rand], [b for _, b in rand])   # :123


def kate_division_mod(m: int, a: Sequence[int], b: int) -> List[int]:
    """arithmetic.rs:322-341 with the modulus given directly."""
    nb = (-b) % m
    q = [0] * (len(a) - 1)
    tmp = 0
    for qi, coeff in zip(range(len(q) - 1, -1, -1), reversed(list(a))):
        lead = (coeff - tmp) % m
        q[qi] = lead
        tmp = lead * nb % m
    return q


def multiopen_verify_proof(k: int, transcript, queries, msm: MSM) -> Guard:
    """poly/multiopen/verifier.rs:14-140.  `msm`: the (usually empty) MSM the opened commitment is accumulated into."""
    c = msm.c
    r = c.r
    x1 = transcript.squeeze_challenge()                        # :31
    x2 = transcript.squeeze_challenge()                        # :35
    sets = construct_intermediate_sets(queries, prover=False)  # :37-38
    if sets is None:
        raise VerifyError("OpeningError")
    commitment_map, point_sets = sets
    q_commitments = [[MSM(c, msm.g, msm.w, msm.u), 1] for _ in point_sets]   # :42-44 (accumulator, next x_1 power)
    q_eval_sets = [[0] * len(ps) for ps in point_sets]         # :48-51
    for data in reversed(commitment_map):                      # :54-86: increasing powers of x_1 from the last commitment
        acc = q_commitments[data["set_index"]]
        cm = data["commitment"]
        if isinstance(cm, MSM):                                # :62-66
            scaled = cm.clone()
            scaled.scale(acc[1])
            acc[0].add_msm(scaled)
        else:
            acc[0].append_term(acc[1], cm)                     # :59-61
        es = q_eval_sets[data["set_index"]]
        for i, ev in enumerate(data["evals"]):                 # :68-70
            es[i] = (es[i] + ev * acc[1]) % r
        acc[1] = acc[1] * x1 % r
    try:
        q_prime_commitment = transcript.read_point()           # :90
    except Exception as e:
        raise VerifyError("SamplingError") from e
    x3 = transcript.squeeze_challenge()                        # :94
    u_evals = []
    for _ in q_eval_sets:                                      # :98-101
        try:
            u_evals.append(transcript.read_scalar())
        except Exception as e:
            raise VerifyError("SamplingError") from e
    msm_eval = 0
    for points, evals, proof_eval in zip(point_sets, q_eval_sets, u_evals):   # :105-119
        r_eval = eval_polynomial_mod(r, lagrange_interpolate(r, points, evals), x3)
        ev = (proof_eval - r_eval) % r
        for pt in points:
            ev = ev * inv((x3 - pt) % r, r) % r
        msm_eval = (msm_eval * x2 + ev) % r
    x4 = transcript.squeeze_challenge()                        # :123
    msm.append_term(1, q_prime_commitment)                     # :126
    v = msm_eval
    for (q_commitment, _), q_eval in zip(q_commitments, u_evals):   # :127-134
        msm.scale(x4)
        msm.add_msm(q_commitment)
        v = (v * x4 + q_eval) % r
    return ipa_verify_proof(k, msm, transcript, x3, v)         # :137
