"""Ad-hoc GPU timing used during development (not part of bench.py's contract)."""
import ctypes
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from halo2_b200 import lib as L  # noqa: E402

lib = L.init()


def mulbench():
    for field in (0, 1, 0x100, 0x101):
        for tpb, blocks in ((128, 148 * 4), (256, 148 * 4), (256, 148 * 8)):
            ms = ctypes.c_float()
            iters = 2000
            L.check(lib.h2_bench_field_mul(field, tpb, blocks, iters, ctypes.byref(ms)))
            muls = tpb * blocks * iters * 4
            print(f"field={field} tpb={tpb} blocks={blocks}: {ms.value:.3f} ms  {muls / ms.value / 1e6:.1f} G mulmod/s", flush=True)


def latbench():
    names = ["mul x1 chain", "mul x2 chains", "mul x4 chains", "xyzz_double", "xyzz_add", "xyzz_add_mixed", "fe_mul2 pair"]
    per = [1, 2, 4, 1, 1, 1, 2]
    for mode in range(7):
        ms = ctypes.c_float()
        iters = 2000
        L.check(lib.h2_bench_latency(mode, iters, ctypes.byref(ms)))
        print(f"latency {names[mode]:16s}: {ms.value * 1e3 / iters:.3f} us/iter  ({ms.value * 1e3 / iters / per[mode]:.3f} us per op)", flush=True)


def rand_scalars(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randint(0, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=g)
    x[:, 7] &= 0x3FFFFFFF  # < 2^254 < modulus: canonical
    return x


def ntt_time(log_n, field=0, reps=10):
    n = 1 << log_n
    a = rand_scalars(n, 1)
    out = torch.empty_like(a)
    omega = np.frombuffer((5).to_bytes(32, "little"), dtype=np.uint8).copy()
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        L.check(lib.h2_ntt_dev(field, ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(out.data_ptr()), L.ptr(omega), 0, log_n, ctypes.c_void_p(s)))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        L.check(lib.h2_ntt_dev(field, ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(out.data_ptr()), L.ptr(omega), 0, log_n, ctypes.c_void_p(s)))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"ntt 2^{log_n}: {ms:.3f} ms  {n / ms / 1e6:.2f} G elem/s", flush=True)


def msm_time(log_n, curve=0, c=0, reps=5, extra=0):
    n = (1 << log_n) + extra
    sc = rand_scalars(n, 2)
    bases = torch.empty((n, 16), dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    L.check(lib.h2_dev_gen_points(curve, 7, 0, ctypes.c_size_t(n), ctypes.c_void_p(bases.data_ptr()), ctypes.c_void_p(s)))
    out = torch.empty(24, dtype=torch.int32, device="cuda")
    for _ in range(2):
        L.check(lib.h2_msm_dev(curve, ctypes.c_void_p(sc.data_ptr()), 0, ctypes.c_void_p(bases.data_ptr()), ctypes.c_size_t(n), c, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(s)))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        L.check(lib.h2_msm_dev(curve, ctypes.c_void_p(sc.data_ptr()), 0, ctypes.c_void_p(bases.data_ptr()), ctypes.c_size_t(n), c, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(s)))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"msm n=2^{log_n}+{extra} c={c}: {ms:.3f} ms  {n / ms / 1e3:.2f} M pairs/s", flush=True)


def commit_time(k, curve=1, reps=20):
    """Host-buffer commits against resident bases (the Params::commit path), with / without the window table."""
    n = (1 << k) + 1
    bases = torch.empty((n, 16), dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    L.check(lib.h2_dev_gen_points(curve, 11, 0, ctypes.c_size_t(n), ctypes.c_void_p(bases.data_ptr()), ctypes.c_void_p(s)))
    torch.cuda.synchronize()
    hb = bases.cpu().numpy().view(np.uint8).reshape(n, 64)
    sc = rand_scalars(n - 1, 3).cpu().numpy().view(np.uint8).reshape(n - 1, 32)
    blind = np.frombuffer((7).to_bytes(32, "little"), dtype=np.uint8).copy()
    for flags in (0, 1):
        h = ctypes.c_uint64(0)
        t0 = time.time()
        L.check(lib.h2_bases_register_ex(curve, L.ptr(hb), ctypes.c_size_t(n), 1, 0, flags, ctypes.byref(h)))
        treg = time.time() - t0
        out = np.zeros(96, dtype=np.uint8)
        for _ in range(3):
            L.check(lib.h2_msm_registered(h, L.ptr(sc), ctypes.c_size_t(n - 1), L.ptr(blind), 0, L.ptr(out)))
        t0 = time.time()
        for _ in range(reps):
            L.check(lib.h2_msm_registered(h, L.ptr(sc), ctypes.c_size_t(n - 1), L.ptr(blind), 0, L.ptr(out)))
        dt = (time.time() - t0) / reps
        print(f"commit k={k} table={flags}: {dt * 1e3:.3f} ms per commit (register {treg * 1e3:.1f} ms)", flush=True)
        L.check(lib.h2_bases_release(h))


if __name__ == "__main__":
    what = sys.argv[1:] or ["mul", "ntt", "msm"]
    if "mul" in what:
        mulbench()
    if "lat" in what:
        latbench()
    if "commit" in what:
        for k in (10, 14, 16, 20):
            commit_time(k)
    if "ntt" in what:
        for k in (14, 16, 20, 24):
            ntt_time(k)
    if "msm" in what:
        msm_time(14, extra=1)
        msm_time(16)
        for c in (0, 13, 15, 17):
            msm_time(20, c=c)
