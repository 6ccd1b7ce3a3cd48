"""Ad-hoc: device MSM time at 2^k for the skewed scalar distributions provers produce (exact-sort fallback, split buckets)."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from halo2_b200 import lib as L  # noqa: E402

lib = L.init()
k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << k
s = torch.cuda.current_stream().cuda_stream
bases = torch.empty((n, 16), dtype=torch.int32, device="cuda")
L.check(lib.h2_dev_gen_points(0, 7, 0, ctypes.c_size_t(n), ctypes.c_void_p(bases.data_ptr()), ctypes.c_void_p(s)))
g = torch.Generator(device="cuda").manual_seed(3)
uni = torch.randint(0, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=g)
uni[:, 7] &= 0x3FFFFFFF
cases = {"uniform": uni}
small = torch.zeros_like(uni); small[:, 0] = torch.randint(0, 1000, (n,), dtype=torch.int32, device="cuda", generator=g); cases["small < 1000"] = small
bits = torch.zeros_like(uni); bits[:, 0] = torch.randint(0, 2, (n,), dtype=torch.int32, device="cuda", generator=g); cases["0/1 column"] = bits
eq = uni[:1].repeat(n, 1).contiguous(); cases["all equal"] = eq
half = uni.clone(); half[::2] = 0; cases["50% zeros"] = half
rep = uni[torch.randint(0, 16, (n,), device="cuda", generator=g)].contiguous(); cases["16 distinct values"] = rep
u64 = torch.zeros_like(uni); u64[:, :2] = uni[:, :2]; cases["64-bit values"] = u64
out = torch.empty(24, dtype=torch.int32, device="cuda")
for name, sc in cases.items():
    for _ in range(2):
        L.check(lib.h2_msm_dev(0, ctypes.c_void_p(sc.data_ptr()), 0, ctypes.c_void_p(bases.data_ptr()), ctypes.c_size_t(n), 0, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(s)))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        L.check(lib.h2_msm_dev(0, ctypes.c_void_p(sc.data_ptr()), 0, ctypes.c_void_p(bases.data_ptr()), ctypes.c_size_t(n), 0, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(s)))
    e1.record(); torch.cuda.synchronize()
    f = ctypes.c_uint32(0); L.check(lib.h2_test_last_msm_flags(ctypes.byref(f)))
    print(f"2^{k} {name:20s}: {e0.elapsed_time(e1) / 3:8.3f} ms   flags={f.value} (1 = split buckets, 2 = exact sort)", flush=True)
