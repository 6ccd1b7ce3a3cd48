"""The fixed columns and the permutation polynomials of the reference's `plonk_api` test circuit
(/root/reference/halo2_proofs/tests/plonk_api.rs:21-420), rebuilt from its source so that the golden
commitments it pins (:958-982) can be recomputed: every one of them is
`params.commit_lagrange(column, Blind::default())` with `params = Params::<EqAffine>::new(5)`
(plonk/keygen.rs:233-236, plonk/permutation/keygen.rs:135-150) -- i.e. hash_to_curve -> EC-iFFT -> MSM.

Test infrastructure only.  Pure integer bookkeeping; the caller supplies omega / delta / zeta.

Layout (SimpleFloorPlanner, circuit/floor_planner/single_pass.rs: a region starts at the first row free in all of
its columns): row 0 = public_input (sp = 1); then ten times a raw_multiply row (sa = sb = 0, sc = sm = 1) followed by
a raw_add row (sa = sb = sc = 1, sm = 0): rows 1..20.  Fixed columns in creation order (plonk_api.rs:293-307):
sf, sm, sa, sb, sc, sp, sl.  The table column sl holds [instance, a, a, 0] and is then filled with its row-0 value
up to the last usable row (circuit/table_layouter.rs:96, plonk/keygen.rs:152-173), usable rows = n - (blinding_factors
+ 1) with blinding_factors = max(3, 1) + 2 = 5 (plonk/circuit.rs:1435-1460).
Equality columns in enable_equality order (plonk_api.rs:299-301, :348-356): a, b, c, sf, e, d, p, sm, sa, sb, sc, sp;
per iteration copy(a0, a1) and copy(b1, c0) (plonk_api.rs:399-400), merged as plonk/permutation/keygen.rs:44-100.
"""
K = 5
N = 1 << K
BLINDING_FACTORS = 5
USABLE_ROWS = N - (BLINDING_FACTORS + 1)
A_SMALL = 2834758237  # plonk_api.rs:421: a = Fp::from(2834758237) * Fp::ZETA


def fixed_columns(modulus: int, zeta: int):
    instance = 2
    a = A_SMALL * zeta % modulus
    mul_rows = [1 + 2 * i for i in range(10)]
    add_rows = [2 + 2 * i for i in range(10)]
    col = lambda rows: [1 if r in rows else 0 for r in range(N)]
    sf = [0] * N
    sm = col(mul_rows)
    sa = col(add_rows)
    sb = col(add_rows)
    sc = col(mul_rows + add_rows)
    sp = col([0])
    table = [instance, a, a, 0]
    sl = [0] * N
    for r in range(USABLE_ROWS):
        sl[r] = table[r] if r < len(table) else table[0]
    return [sf, sm, sa, sb, sc, sp, sl]


def permutation_columns(modulus: int, omega: int, delta: int):
    ncols = 12
    COL_A, COL_B, COL_C = 0, 1, 2
    mapping = [[(i, j) for j in range(N)] for i in range(ncols)]
    aux = [[(i, j) for j in range(N)] for i in range(ncols)]
    sizes = [[1] * N for _ in range(ncols)]

    def copy(lc, lr, rc, rr):
        left, right = aux[lc][lr], aux[rc][rr]
        if left == right:
            return
        if sizes[left[0]][left[1]] < sizes[right[0]][right[1]]:
            left, right = right, left
        sizes[left[0]][left[1]] += sizes[right[0]][right[1]]
        i = right
        while True:
            aux[i[0]][i[1]] = left
            i = mapping[i[0]][i[1]]
            if i == right:
                break
        mapping[lc][lr], mapping[rc][rr] = mapping[rc][rr], mapping[lc][lr]

    for it in range(10):
        rm, ra = 1 + 2 * it, 2 + 2 * it
        for _ in range(2):                      # StandardCs::copy constrains twice (plonk_api.rs:216-217)
            copy(COL_A, rm, COL_A, ra)          # copy(a0, a1)
        for _ in range(2):
            copy(COL_B, ra, COL_C, rm)          # copy(b1, c0)

    omega_powers = [pow(omega, j, modulus) for j in range(N)]
    deltas = [pow(delta, i, modulus) for i in range(ncols)]
    return [[deltas[mapping[i][j][0]] * omega_powers[mapping[i][j][1]] % modulus for j in range(N)]
            for i in range(ncols)]
