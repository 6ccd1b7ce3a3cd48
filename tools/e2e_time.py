"""Ad-hoc: h2_msm (pinned host buffers) wall time at 2^k for chunked / unchunked upload and both encodings."""
import ctypes
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from halo2_b200 import lib as L  # noqa: E402

lib = L.init()
k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << k
g = torch.Generator(device="cuda").manual_seed(1)
sc = torch.randint(0, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=g)
sc[:, 7] &= 0x3FFFFFFF
bases = torch.empty((n, 16), dtype=torch.int32, device="cuda")
L.check(lib.h2_dev_gen_points(0, 7, 0, ctypes.c_size_t(n), ctypes.c_void_p(bases.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
torch.cuda.synchronize()
sc_h = sc.cpu().pin_memory()
bs_mont = bases.cpu().pin_memory()          # h2_dev_gen_points emits Montgomery coordinates
L.check(lib.h2_dev_convert(0, ctypes.c_void_p(bases.data_ptr()), ctypes.c_size_t(2 * n), 0, None))
torch.cuda.synchronize()
bs_canon = bases.cpu().pin_memory()
out = np.zeros(96, dtype=np.uint8)
# raw link speed
dst = torch.empty_like(bases)
torch.cuda.synchronize(); t0 = time.time()
for _ in range(5):
    dst.copy_(bs_canon, non_blocking=True)
torch.cuda.synchronize(); dt = (time.time() - t0) / 5
print(f"H2D 64 MiB: {dt * 1e3:.3f} ms  {bs_canon.numel() * 4 / dt / 1e9:.1f} GB/s", flush=True)
for thr, name in ((40, "one upload"), (k, "2 chunks"), (k - 1, "3 chunks"), (k - 3, "4 chunks")):
    L.check(lib.h2_test_set_chunk_threshold(thr))
    for repr_, bs, rname in ((0, bs_canon, "canonical"), (1, bs_mont, "montgomery")):
        for _ in range(3):
            L.check(lib.h2_msm(0, ctypes.c_void_p(sc_h.data_ptr()), ctypes.c_void_p(bs.data_ptr()), ctypes.c_size_t(n), repr_, L.ptr(out)))
        t0 = time.time()
        reps = 10
        for _ in range(reps):
            L.check(lib.h2_msm(0, ctypes.c_void_p(sc_h.data_ptr()), ctypes.c_void_p(bs.data_ptr()), ctypes.c_size_t(n), repr_, L.ptr(out)))
        dt = (time.time() - t0) / reps
        print(f"h2_msm 2^{k} {name:10s} {rname:10s}: {dt * 1e3:.3f} ms", flush=True)
