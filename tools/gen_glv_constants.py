"""Derives the GLV constants of csrc/glv.cuh from first principles and self-checks them.

Pallas / Vesta have j-invariant 0: phi(x, y) = (zeta x, y) acts as multiplication by lambda, with
zeta^3 = 1 in the base field and lambda^2 + lambda + 1 = 0 mod r.  (a1, b1), (a2, b2) is a reduced
basis of the lattice {(a, b): a + b lambda = 0 mod r} (extended Euclid on (r, lambda)); a scalar k
splits as k = k1 + k2 lambda with k1 = k - c1 a1 - c2 a2, k2 = -c1 b1 - c2 b2, c1 = round(k b2 / r),
c2 = round(-k b1 / r), |k1|, |k2| < 2^127.  The device uses g_i = round(2^384 |b_j| / r) and
c_i = floor((k g_i + 2^383) / 2^384): within 1/2 + 2^-130 of the real quotient, so
|k1| <= (1/2 + eps)(a1 + a2), |k2| <= (1/2 + eps)(|b1| + b2) < 0.867 x 2^127 (asserted below).  The identity k1 + k2 lambda = k (mod r) holds for ANY integers c1, c2 because the basis vectors
are in the lattice, so correctness never depends on the rounding.
"""
import sys
from math import isqrt

sys.path.insert(0, ".")
from oracle import pasta as o  # noqa: E402  (build-time tool, not part of the product)


def find_lambda_zeta(c):
    r, p = c.r, c.p
    lam = pow(5, (r - 1) // 3, r)
    g = o.generator(c)
    for l in (lam, lam * lam % r):
        q = o.to_affine(c, o.scalar_mul(c, l, g))
        for z in o.zeta_candidates(c.base):
            if q == (z * g[0] % p, g[1]):
                return l, z
    raise SystemExit("no (lambda, zeta) pair")


def lattice(r, lam):
    r0, r1, t0, t1 = r, lam, 0, 1
    rows = []
    while r1:
        q = r0 // r1
        r0, r1, t0, t1 = r1, r0 - q * r1, t1, t0 - q * t1
        rows.append((r0, t0))
    sq = isqrt(r)
    for i, (ri, ti) in enumerate(rows):
        if ri < sq:
            v1 = (ri, -ti)
            c1, c2 = (rows[i - 1][0], -rows[i - 1][1]), (rows[i + 1][0], -rows[i + 1][1])
            v2 = c1 if c1[0] ** 2 + c1[1] ** 2 <= c2[0] ** 2 + c2[1] ** 2 else c2
            return v1, v2
    raise SystemExit("no basis")


def limbs(x, n):
    return ", ".join("0x%08xu" % ((x >> (32 * i)) & 0xFFFFFFFF) for i in range(n))


def main():
    out = []
    for c, tag in ((o.PALLAS, "FpParams"), (o.VESTA, "FqParams")):
        r, p = c.r, c.p
        lam, zeta = find_lambda_zeta(c)
        (a1, b1), (a2, b2) = lattice(r, lam)
        assert (a1 + b1 * lam) % r == 0 and (a2 + b2 * lam) % r == 0 and a1 * b2 - a2 * b1 == r
        assert a1 > 0 and a2 > 0 and b1 < 0 and b2 > 0
        assert (a1 + a2) * 1001 // 2000 < 1 << 127 and (b2 - b1) * 1001 // 2000 < 1 << 127   # |k_i| < 2^127 rigorously
        g1 = ((b2 << 384) + r // 2) // r       # c1 ~ k b2 / r
        g2 = (((-b1) << 384) + r // 2) // r    # c2 ~ k |b1| / r
        # exhaustive-ish self check of the device formula
        import random
        random.seed(7)
        mx = 0
        for k in [0, 1, 2, r - 1, r - 2, lam, r - lam, (1 << 254), (1 << 254) - 1] + [random.randrange(r) for _ in range(20000)]:
            c1, c2 = (k * g1 + (1 << 383)) >> 384, (k * g2 + (1 << 383)) >> 384
            k1 = k - c1 * a1 - c2 * a2
            k2 = c1 * (-b1) - c2 * b2
            assert (k1 + k2 * lam - k) % r == 0
            mx = max(mx, abs(k1).bit_length(), abs(k2).bit_length())
        assert mx <= 127, mx
        R = (1 << 256) % p
        out.append((c.name, tag, lam, zeta, zeta * R % p, a1, -b1, a2, b2, g1, g2, mx))
    for name, tag, lam, zeta, zeta_m, a1, nb1, a2, b2, g1, g2, mx in out:
        print(f"// {name}: lambda = {hex(lam)}")
        print(f"//          zeta = {hex(zeta)}   (max |k_i| bits observed: {mx})")
        print(f"template <> struct GlvConst<{tag}> {{")
        print(f"    static H2_HD uint32_t zeta_mont(int i) {{ constexpr uint32_t v[8] = {{{limbs(zeta_m, 8)}}}; return v[i]; }}")
        for nm, val in (("a1", a1), ("nb1", nb1), ("a2", a2), ("b2", b2)):
            print(f"    static H2_HD uint32_t {nm}(int i) {{ constexpr uint32_t v[4] = {{{limbs(val, 4)}}}; return v[i]; }}")
        for nm, val in (("g1", g1), ("g2", g2)):
            assert val < (1 << 288)
            print(f"    static H2_HD uint32_t {nm}(int i) {{ constexpr uint32_t v[9] = {{{limbs(val, 9)}}}; return v[i]; }}")
        print("};")


if __name__ == "__main__":
    main()
