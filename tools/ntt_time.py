"""best_fft 2^log_n device-resident timing (CUDA events), both fields."""
import ctypes, sys
sys.path.insert(0, ".")
import torch
from halo2_b200 import lib as L
lib = L.init()
dev = torch.device("cuda", 0)
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = {"fp": 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001, "fq": 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001}
for log_n in [int(a) for a in sys.argv[1:]] or [20]:
  n = 1 << log_n
  for tma in (1, 0):
    L.check(lib.h2_test_set_ntt_tma(tma))
    for f in ("fp", "fq"):
        m = P[f]
        w = pow(5, (m - 1) >> 32, m)
        for _ in range(log_n, 32):
            w = w * w % m
        ob = L.fe_bytes(w)
        g = torch.Generator(device=dev).manual_seed(1)
        bufs = [torch.randint(-2**31, 2**31 - 1, (n, 8), dtype=torch.int32, device=dev, generator=g) for _ in range(5)]
        for b in bufs:
            b[:, 7] &= 0x3FFFFFFF
        outs = [torch.empty_like(bufs[0]) for _ in range(5)]
        def step(i):
            L.check(lib.h2_ntt_dev(L.FIELD_ID[f], ctypes.c_void_p(bufs[i % 5].data_ptr()), ctypes.c_void_p(outs[i % 5].data_ptr()), L.ptr(ob), L.REPR_CANONICAL, log_n, sp))
        for i in range(3):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"ntt 2^{log_n} {f} tma={tma}: {ms:.4f} ms  {n / ms / 1e6:.2f} G elems/s")
