"""Size-independent properties at (and above) BASELINE.json's full sizes, where the CPU oracle is too
slow to be the checker: different window sizes / GLV on-off / sharded-and-summed must give the same
point; NTT followed by the inverse NTT must give the input back; linearity of the transform."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import cref, pasta  # noqa: E402

SEED = 0x48414C4F32


@pytest.fixture(scope="module")
def dev():
    import torch
    from halo2_b200 import lib as L
    lib = L.init()
    return torch, L, lib


def _rand_scalars(torch, n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randint(-2**31, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=g)
    x[:, 7] &= 0x3FFFFFFF
    return x


def _msm_dev(torch, L, lib, curve, sc, bases, c=0):
    out = torch.zeros(24, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    L.check(lib.h2_msm_dev(L.CURVE_ID[curve], ctypes.c_void_p(sc.data_ptr()), L.REPR_CANONICAL, ctypes.c_void_p(bases.data_ptr()),
                           ctypes.c_size_t(sc.shape[0]), c, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(s)))
    torch.cuda.synchronize()
    xyz = out.cpu().numpy().view(np.uint8).reshape(3, 32)
    m = pasta.CURVES[curve].p
    rinv = pow((1 << 256) % m, m - 2, m)
    canon = cref.ints_to_bytes([v * rinv % m for v in cref.bytes_to_ints(xyz)]).reshape(-1)
    return cref.bytes_to_affine(cref.jac_to_affine(curve, canon))


@pytest.mark.parametrize("curve,log_n", [("pallas", 22), ("vesta", 20), ("vesta", 24)])   # 2^24: default window 19, 9 k-entry bins
def test_msm_full_size_consistency(dev, curve, log_n):
    torch, L, lib = dev
    n = 1 << log_n
    cid = L.CURVE_ID[curve]
    sc = _rand_scalars(torch, n, SEED + log_n)
    bases = torch.empty((n, 16), dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    L.check(lib.h2_dev_gen_points(cid, SEED + 5, 0, ctypes.c_size_t(n), ctypes.c_void_p(bases.data_ptr()), ctypes.c_void_p(s)))
    try:
        ref = _msm_dev(torch, L, lib, curve, sc, bases)                      # GLV, c = 16
        assert ref is not None
        assert _msm_dev(torch, L, lib, curve, sc, bases, 13) == ref           # another window size
        L.check(lib.h2_set_glv(0))
        assert _msm_dev(torch, L, lib, curve, sc, bases) == ref               # plain 255-bit path
        assert _msm_dev(torch, L, lib, curve, sc, bases, 19) == ref
        L.check(lib.h2_set_glv(1))
        # sharded (the multi-GPU decomposition on one device): 4 contiguous shards, partials summed
        parts = np.zeros((4, 96), dtype=np.uint8)
        q = n // 4
        for k in range(4):
            out = torch.zeros(24, dtype=torch.int32, device="cuda")
            L.check(lib.h2_msm_dev(cid, ctypes.c_void_p(sc[k * q:(k + 1) * q].data_ptr()), L.REPR_CANONICAL,
                                   ctypes.c_void_p(bases[k * q:(k + 1) * q].data_ptr()), ctypes.c_size_t(q), 0,
                                   ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(s)))
            torch.cuda.synchronize()
            parts[k] = out.cpu().numpy().view(np.uint8)
        tot = np.zeros(96, dtype=np.uint8)
        L.check(lib.h2_point_sum(cid, L.ptr(parts), ctypes.c_size_t(4), L.REPR_MONTGOMERY, L.ptr(tot)))
        m = pasta.CURVES[curve].p
        rinv = pow((1 << 256) % m, m - 2, m)
        canon = cref.ints_to_bytes([v * rinv % m for v in cref.bytes_to_ints(tot.reshape(3, 32))]).reshape(-1)
        assert cref.bytes_to_affine(cref.jac_to_affine(curve, canon)) == ref
        # the whole problem against the oracle (the reference algorithm restated in C, all host threads): ~0.6 s at 2^20,
        # ~3 s at 2^22, ~12 s at 2^24 on the GPU box
        pb_all = bases.clone()
        L.check(lib.h2_dev_convert(L.FIELD_ID[L.BASE_FIELD[curve]], ctypes.c_void_p(pb_all.data_ptr()), ctypes.c_size_t(2 * n), 0, ctypes.c_void_p(s)))
        torch.cuda.synchronize()
        pb_host = pb_all.cpu().numpy().view(np.uint8).reshape(n, 64)
        del pb_all
        kb_host = sc.cpu().numpy().view(np.uint8).reshape(n, 32)
        assert ref == cref.bytes_to_affine(cref.best_multiexp(curve, kb_host, pb_host)), "full-size MSM differs from the oracle"
        del pb_host, kb_host
        # spot check against the oracle on a prefix (same bases): 2^12 terms
        k = 1 << 12
        pb = bases[:k].clone()
        L.check(lib.h2_dev_convert(L.FIELD_ID[L.BASE_FIELD[curve]], ctypes.c_void_p(pb.data_ptr()), ctypes.c_size_t(2 * k), 0, ctypes.c_void_p(s)))
        torch.cuda.synchronize()
        pbh = pb.cpu().numpy().view(np.uint8).reshape(k, 64)
        kbh = sc[:k].cpu().numpy().view(np.uint8).reshape(k, 32)
        assert _msm_dev(torch, L, lib, curve, sc[:k].contiguous(), bases[:k].contiguous()) == cref.bytes_to_affine(cref.best_multiexp(curve, kbh, pbh))
    finally:
        lib.h2_set_glv(1)


@pytest.mark.parametrize("field,log_n", [("fp", 24), ("fq", 20)])
def test_ntt_full_size_round_trip_and_linearity(dev, field, log_n):
    torch, L, lib = dev
    n = 1 << log_n
    m = pasta.FIELDS[field]
    fid = L.FIELD_ID[field]
    s = torch.cuda.current_stream().cuda_stream
    w = pasta.omega_for_k(field, log_n)
    w_inv = pow(w, m - 2, m)
    a = _rand_scalars(torch, n, SEED + 3)
    b = _rand_scalars(torch, n, SEED + 4)

    def ntt(x, omega):
        out = torch.empty_like(x)
        L.check(lib.h2_ntt_dev(fid, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()), L.ptr(L.fe_bytes(omega)),
                               L.REPR_CANONICAL, log_n, ctypes.c_void_p(s)))
        return out

    # values are interpreted as Montgomery residues: any 255-bit pattern < m is a valid element
    fa = ntt(a, w)
    back = ntt(fa, w_inv)                 # = n * a  (no 1/n scaling in best_fft)
    torch.cuda.synchronize()
    # the whole transform against the oracle's butterfly network (the data are raw residues and the network is linear, so
    # the raw outputs are best_fft of the raw inputs)
    ah = a.cpu().numpy().view(np.uint8).reshape(n, 32)
    assert (fa.cpu().numpy().view(np.uint8).reshape(n, 32) == cref.best_fft(field, ah, w, log_n)).all(), "full-size NTT differs from the oracle"
    # compare n * a with back on a sample of positions using host big ints (Montgomery form is linear)
    idx = [0, 1, 2, n // 3, n // 2 + 7, n - 1]
    bh = back.cpu().numpy().view(np.uint8).reshape(n, 32)
    for i in idx:
        av = int.from_bytes(ah[i].tobytes(), "little")
        assert int.from_bytes(bh[i].tobytes(), "little") == av * n % m, i
    # linearity against the closed form: b' = delta at position 5  ->  ntt(b')[p] = w^(5 p) * b'_5
    # (adding whole 2^24-element vectors mod m on the host would be too slow in Python)
    sel = torch.tensor(idx, device="cuda")
    d = torch.zeros_like(a)
    d[5] = b[5]
    fd = ntt(d, w)
    torch.cuda.synchronize()
    fdh = fd[sel].cpu().numpy().view(np.uint8).reshape(len(idx), 32)
    b5 = int.from_bytes(b[5].cpu().numpy().view(np.uint8).tobytes(), "little")
    for row, p in zip(fdh, idx):
        # Montgomery residues: out = b5 * w^(5p) as field elements => residue(out) = residue(b5) * w^(5p)
        assert int.from_bytes(row.tobytes(), "little") == b5 * pow(w, 5 * p, m) % m, p
    # ... and additivity on the delta: ntt(a + b') = ntt(a) + ntt(b') at the sampled positions
    a2 = a.clone()
    a5 = int.from_bytes(a[5].cpu().numpy().view(np.uint8).tobytes(), "little")
    s5 = (a5 + b5) % m
    a2[5] = torch.tensor(list(np.frombuffer(s5.to_bytes(32, "little"), dtype=np.int32)), dtype=torch.int32, device="cuda")
    fa2 = ntt(a2, w)
    torch.cuda.synchronize()
    fah = fa[sel].cpu().numpy().view(np.uint8).reshape(len(idx), 32)
    fa2h = fa2[sel].cpu().numpy().view(np.uint8).reshape(len(idx), 32)
    for x, y, z in zip(fah, fdh, fa2h):
        assert (int.from_bytes(x.tobytes(), "little") + int.from_bytes(y.tobytes(), "little")) % m == int.from_bytes(z.tobytes(), "little")


def test_ec_fft_round_trip_and_commit_2pow13():
    """best_fft at G = curve point at k = 13 (4x the oracle-checked size; both butterfly forms: quads up to log n = 12, one
    thread per butterfly above): the inverse transform with the 2^-k scaling brings the generators back, and the reference's
    own property poly/commitment.rs:258-302 holds for the derived g_lagrange: commit_lagrange(a) == commit(lagrange_to_coeff(a))."""
    import halo2_b200 as h2
    curve, c = "vesta", pasta.VESTA
    for k in (12, 13):
        n, r = 1 << k, c.r
        g = cref.gen_points(curve, SEED + 950 + k, n + 1)
        w = pasta.omega_for_k(c.scalar, k)
        jac = cref.affine_to_jacobian_bytes(g[:n])
        fwd = h2.best_fft_curve(jac.copy(), w, k, curve)
        from halo2_b200 import lib as L
        L.check(L.init().h2_ec_fft(L.CURVE_ID[curve], L.ptr(fwd), L.ptr(L.fe_bytes(pasta.inv(w, r))), ctypes.c_uint32(k),
                                   L.ptr(L.fe_bytes(pow(pasta.inv(2, r), k, r))), L.REPR_CANONICAL))
        assert (h2.batch_normalize(fwd, curve) == g[:n]).all(), k
        params = h2.Params.from_generators(curve, k, g[:n], g[n:])
        dom = h2.EvaluationDomain(c.scalar, 2, k, pasta.zeta_candidates(c.scalar)[0])
        a = cref.gen_scalars(c.scalar, SEED + 960 + k, n)
        lhs = h2.batch_normalize(params.commit_lagrange(a, h2.Blind(3)).reshape(1, 96), curve)
        rhs = h2.batch_normalize(params.commit(dom.lagrange_to_coeff(a), h2.Blind(3)).reshape(1, 96), curve)
        assert (lhs == rhs).all() and lhs.any(), k
        params.close()
