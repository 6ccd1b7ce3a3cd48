"""Times the g -> g_lagrange derivation (h2_params_lagrange: EC-iFFT + scale + batch_normalize) on the GPU next to the
CPU restatement.  Usage: python tools/ecfft_time.py [k ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_b200  # noqa: E402
from halo2_b200 import lib as L  # noqa: E402
from oracle import cref, pasta  # noqa: E402

ks = [int(a) for a in sys.argv[1:]] or [10, 12, 14]
curve, c = "vesta", pasta.VESTA
L.init()
for k in ks:
    g = cref.gen_points(curve, 1000 + k, 1 << k)
    reps = 3
    line = f"k={k}:"
    for form in (0, 1):
        L.check(L.init().h2_test_set_ecfft_quad(form))
        halo2_b200.lagrange_generators(curve, k, g)
        t = time.perf_counter()
        for _ in range(reps):
            out = halo2_b200.lagrange_generators(curve, k, g)
        gpu_ms = (time.perf_counter() - t) / reps * 1e3
        line += f" GPU {'quad' if form else 'thread'} form {gpu_ms:.2f} ms,"
    L.check(L.init().h2_test_set_ecfft_quad(-1))
    if k <= int(os.environ.get("ECFFT_CPU_MAX_K", "12")):
        r = c.r
        t = time.perf_counter()
        want = cref.params_lagrange(curve, g, k, pasta.inv(pasta.omega_for_k(c.scalar, k), r), pow(pasta.inv(2, r), k, r))
        cpu_ms = (time.perf_counter() - t) * 1e3
        line += f" CPU restatement ({cref.default_threads()} threads) {cpu_ms:.1f} ms, same result: {bool((out == want).all())}"
    print(line, flush=True)
