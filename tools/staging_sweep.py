"""Ad-hoc: pageable-memory rates through the staging ring by copy-thread count and NT stores: raw upload (h2_poly_upload, 96 MiB)
and h2_msm 2^20 end to end (the bench's headline e2e)."""
import ctypes, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import halo2_b200 as h2
from halo2_b200 import lib as L

lib = L.init()
n = 3 << 20
p = h2.ResidentPoly("fp", n)
pageable = np.ones((n, 32), dtype=np.uint8)
k = 20
m = 1 << k
g = torch.Generator(device="cuda").manual_seed(1)
sc = torch.randint(0, 2**31 - 1, (m, 8), dtype=torch.int32, device="cuda", generator=g)
sc[:, 7] &= 0x3FFFFFFF
bases = torch.empty((m, 16), dtype=torch.int32, device="cuda")
L.check(lib.h2_dev_gen_points(0, 7, 0, ctypes.c_size_t(m), ctypes.c_void_p(bases.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
L.check(lib.h2_dev_convert(0, ctypes.c_void_p(bases.data_ptr()), ctypes.c_size_t(2 * m), 0, None))
torch.cuda.synchronize()
sc_pg = [sc.cpu().numpy().copy() for _ in range(2)]
bs_pg = [bases.cpu().numpy().copy() for _ in range(2)]
sc_pin, bs_pin = sc.cpu().pin_memory(), bases.cpu().pin_memory()
out = np.zeros(96, dtype=np.uint8)


def up(reps=5):
    L.check(lib.h2_poly_upload(p._h, L.ptr(pageable), ctypes.c_size_t(n), L.REPR_MONTGOMERY))
    t0 = time.time()
    for _ in range(reps):
        L.check(lib.h2_poly_upload(p._h, L.ptr(pageable), ctypes.c_size_t(n), L.REPR_MONTGOMERY))
    dt = (time.time() - t0) / reps
    return n * 32 / dt / 1e9


def msm(pinned=False, reps=8):
    def one(i):
        if pinned:
            L.check(lib.h2_msm(0, ctypes.c_void_p(sc_pin.data_ptr()), ctypes.c_void_p(bs_pin.data_ptr()), ctypes.c_size_t(m), 0, L.ptr(out)))
        else:
            L.check(lib.h2_msm(0, L.ptr(sc_pg[i % 2]), L.ptr(bs_pg[i % 2]), ctypes.c_size_t(m), 0, L.ptr(out)))
    for i in range(2):
        one(i)
    t0 = time.time()
    for i in range(reps):
        one(i)
    return (time.time() - t0) / reps * 1e3


print(f"h2_msm 2^20 pinned: {msm(True):.3f} ms", flush=True)
for nt in ((1, 0) if len(sys.argv) < 2 else ()):
    for th in (0, 3, 7, 11, 15, 23, 31):
        L.check(lib.h2_test_set_copy_threads(th, nt))
        print(f"nt={nt} threads={th:2d}: upload {up():5.1f} GB/s   h2_msm 2^20 pageable {msm():.3f} ms", flush=True)

if len(sys.argv) > 1 and sys.argv[1] == "chunks":
    L.check(lib.h2_test_set_copy_threads(15, 1))
    for thr, name in ((40, "1 upload"), (20, "2 chunks"), (19, "3 chunks (default)"), (17, "4 chunks")):
        L.check(lib.h2_test_set_chunk_threshold(thr))
        print(f"{name:20s}: pageable {msm():.3f} ms   pinned {msm(True):.3f} ms", flush=True)
    L.check(lib.h2_test_set_chunk_threshold(19))

if len(sys.argv) > 1 and sys.argv[1] == "cuts":
    L.check(lib.h2_test_set_copy_threads(15, 1))
    for k, thr, cuts in ((3, 19, (2, 8, 16)), (3, 19, (3, 8, 16)), (3, 19, (3, 9, 16)), (3, 19, (4, 9, 16)), (3, 19, (4, 10, 16)), (3, 19, (5, 10, 16)),
                         (4, 17, (1, 4, 10)), (4, 17, (2, 6, 11)), (4, 17, (3, 7, 11)), (4, 17, (3, 6, 10)), (2, 20, (4, 16, 16)), (2, 20, (6, 16, 16))):
        L.check(lib.h2_test_set_chunk_threshold(thr))
        L.check(lib.h2_test_set_chunk_cuts(k, *cuts))
        print(f"k={k} cuts={cuts}: pageable {msm(reps=12):.3f} ms   pinned {msm(True, reps=12):.3f} ms", flush=True)
