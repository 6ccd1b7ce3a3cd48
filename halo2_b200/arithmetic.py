"""Host-side mirror of halo2_proofs::arithmetic::{best_multiexp, small_multiexp, best_fft}
(/root/reference/halo2_proofs/src/arithmetic.rs:143-180, :116-136, :192-255) over the C ABI.

Same names, argument meaning and error behaviour as the reference; elements are numpy uint8
arrays in canonical little-endian form (32 B scalars, 64 B affine points, identity = zeros),
i.e. what `to_repr()` / `coordinates()` give on the Rust side.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

from . import lib as _l


def multiexp_window_bits(n: int) -> int:
    """The reference's window choice (arithmetic.rs:146-152); the engine picks its own c
    (msm_default_window in csrc/msm.cuh) -- the result does not depend on it."""
    if n < 4:
        return 1
    if n < 32:
        return 3
    return int(math.ceil(math.log(float(n))))


def best_multiexp(coeffs, bases, curve: str = "vesta", repr: int = _l.REPR_CANONICAL) -> np.ndarray:
    """sum_i coeffs[i] * bases[i] as a Jacobian point (96 bytes x||y||z, z = 0 for the identity).

    Panics (AssertionError) when the lengths differ, like assert_eq! at arithmetic.rs:144."""
    lib = _l.init()
    c = _l.as_u8(coeffs, 32)
    b = _l.as_u8(bases, 64)
    assert c.shape[0] == b.shape[0], "best_multiexp: coeffs.len() != bases.len()"
    out = np.zeros(96, dtype=np.uint8)
    _l.check(lib.h2_msm(_l.CURVE_ID[curve], _l.ptr(c), _l.ptr(b), ctypes.c_size_t(c.shape[0]), int(repr), _l.ptr(out)))
    return out


def best_fft(a, omega, log_n: int, field: str = "fp", repr: int = _l.REPR_CANONICAL) -> np.ndarray:
    """In-place radix-2 network of arithmetic.rs:192-255 on a (2^log_n, 32) uint8 array.

    Panics (AssertionError) when a.len() != 1 << log_n, like assert_eq! at arithmetic.rs:205.
    This is the G = Scalar instantiation; G = curve point (Params::new, poly/commitment.rs:81-82) is
    `best_fft_curve` below -- the Rust shim picks between them on TypeId like the FftGroup bound does."""
    lib = _l.init()
    if not (isinstance(a, np.ndarray) and a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]):
        raise ValueError("best_fft operates in place on a C-contiguous uint8 array")
    arr = a.reshape(-1, 32)
    assert arr.shape[0] == 1 << log_n, "best_fft: a.len() != 1 << log_n"
    _l.check(lib.h2_ntt(_l.FIELD_ID[field], _l.ptr(arr), _l.ptr(_l.fe_bytes(omega)), ctypes.c_uint32(log_n), int(repr)))
    return a


def small_multiexp(coeffs, bases, curve: str = "vesta", repr: int = _l.REPR_CANONICAL) -> np.ndarray:
    """arithmetic.rs:116-136: the reference's shared-doubling double-and-add for a handful of terms.  The same
    group element as best_multiexp on the same inputs, so it runs through the same engine (h2_msm picks a 4-bit
    window below 64 points); the reference itself does not assert equal lengths here (it zips), this mirror
    truncates to the shorter like zip does."""
    c = _l.as_u8(coeffs, 32)
    b = _l.as_u8(bases, 64)
    m = min(c.shape[0], b.shape[0])
    return best_multiexp(c[:m], b[:m], curve=curve, repr=repr)


def best_fft_curve(a, omega, log_n: int, curve: str = "vesta", repr: int = _l.REPR_CANONICAL) -> np.ndarray:
    """best_fft at G = C::Curve (arithmetic.rs:192-255 through FftGroup, :17-27): in place on a (2^log_n, 96) uint8
    array of Jacobian points x||y||z; `omega` is an element of the curve's scalar field.

    Panics (AssertionError) when a.len() != 1 << log_n, like assert_eq! at arithmetic.rs:205."""
    lib = _l.init()
    if not (isinstance(a, np.ndarray) and a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]):
        raise ValueError("best_fft_curve operates in place on a C-contiguous uint8 array")
    arr = a.reshape(-1, 96)
    assert arr.shape[0] == 1 << log_n, "best_fft: a.len() != 1 << log_n"
    _l.check(lib.h2_ec_fft(_l.CURVE_ID[curve], _l.ptr(arr), _l.ptr(_l.fe_bytes(omega)), ctypes.c_uint32(log_n), None, int(repr)))
    return a


def batch_normalize(points_xyz, curve: str = "vesta", repr: int = _l.REPR_CANONICAL) -> np.ndarray:
    """group::Curve::batch_normalize as the prover calls it on its commitments (plonk/prover.rs:99, :311):
    (n, 96) Jacobian -> (n, 64) affine, identity = zeros."""
    lib = _l.init()
    p = _l.as_u8(points_xyz, 96)
    out = np.zeros((p.shape[0], 64), dtype=np.uint8)
    _l.check(lib.h2_batch_normalize(_l.CURVE_ID[curve], _l.ptr(p), ctypes.c_size_t(p.shape[0]), int(repr), _l.ptr(out)))
    return out


def eval_polynomial(poly, point: int, field: str = "fp") -> int:
    """arithmetic.rs:297-303 on a host coefficient vector ((n, 32) uint8 canonical): sum_i poly[i] * point^i."""
    from .poly import ResidentPoly, eval_polynomial_resident
    p = _l.as_u8(poly, 32)
    if p.shape[0] == 0:
        return 0
    r = ResidentPoly(field, p.shape[0], p)
    try:
        return eval_polynomial_resident([r], [point])[0]
    finally:
        r.close()


def compute_inner_product(a, b, field: str = "fp") -> int:
    """arithmetic.rs:308-319; panics (AssertionError) when the lengths differ, like assert_eq! at :311."""
    from .poly import ResidentPoly, inner_product_resident
    x, y = _l.as_u8(a, 32), _l.as_u8(b, 32)
    assert x.shape[0] == y.shape[0], "compute_inner_product: a.len() != b.len()"
    if x.shape[0] == 0:
        return 0
    rx, ry = ResidentPoly(field, x.shape[0], x), ResidentPoly(field, y.shape[0], y)
    try:
        return inner_product_resident([rx], [ry])[0]
    finally:
        rx.close()
        ry.close()


def kate_division(a, b: int, field: str = "fp") -> np.ndarray:
    """arithmetic.rs:322-341: the quotient of a(X) by (X - b) as an (n - 1, 32) uint8 array."""
    from .poly import ResidentPoly, kate_division_resident
    x = _l.as_u8(a, 32)
    assert x.shape[0] >= 1, "kate_division: empty polynomial"
    if x.shape[0] == 1:
        return np.zeros((0, 32), dtype=np.uint8)
    r = ResidentPoly(field, x.shape[0], x)
    try:
        q = kate_division_resident([r], [b])[0]
        try:
            return q.download(x.shape[0] - 1)
        finally:
            q.close()
    finally:
        r.close()
