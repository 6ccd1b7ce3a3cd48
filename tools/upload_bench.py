"""Host -> device transfer rates seen through the library: pinned vs pageable caller memory, staging ring on / off."""
import ctypes, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import halo2_b200 as h2
from halo2_b200 import lib as L

lib = L.init()
n = 3 << 20                                   # 96 MiB of field elements
p = h2.ResidentPoly("fp", n)
pinned = torch.empty((n, 32), dtype=torch.uint8).pin_memory()
pinned.zero_()
pageable = np.zeros((n, 32), dtype=np.uint8)
pageable[:] = 1


def t(src_ptr, reps=5):
    L.check(lib.h2_poly_upload(p._h, src_ptr, ctypes.c_size_t(n), L.REPR_MONTGOMERY))
    t0 = time.time()
    for _ in range(reps):
        L.check(lib.h2_poly_upload(p._h, src_ptr, ctypes.c_size_t(n), L.REPR_MONTGOMERY))
    dt = (time.time() - t0) / reps
    return dt * 1e3, n * 32 / dt / 1e9


print("pinned           %.2f ms  %.1f GB/s" % t(ctypes.c_void_p(pinned.data_ptr())))
for on in (1, 0):
    L.check(lib.h2_test_set_staging(on))
    print("pageable staging=%d %.2f ms  %.1f GB/s" % ((on,) + t(L.ptr(pageable))))
L.check(lib.h2_test_set_staging(1))
out = np.zeros((n, 32), dtype=np.uint8)
for on in (1, 0):
    L.check(lib.h2_test_set_staging(on))
    t0 = time.time()
    for _ in range(3):
        L.check(lib.h2_poly_download(p._h, L.ptr(out), ctypes.c_size_t(n), L.REPR_MONTGOMERY))
    dt = (time.time() - t0) / 3
    print("download pageable staging=%d %.2f ms  %.1f GB/s" % (on, dt * 1e3, n * 32 / dt / 1e9))
