"""Ad-hoc: k=14 commit / batched commit / IPA opening wall time against the fixed-base table's window size."""
import sys
import time

sys.path.insert(0, ".")
import halo2_b200 as h2  # noqa: E402
from oracle import cref  # noqa: E402  (input generation only)

k = int(sys.argv[1]) if len(sys.argv) > 1 else 14
n = 1 << k
g = cref.gen_points("vesta", 1, n + 2)
pp = cref.gen_scalars("fp", 2, n)
polys = [cref.gen_scalars("fp", 10 + i, n) for i in range(4)]
ch = cref.bytes_to_ints(cref.gen_scalars("fp", 3, k))
lr = cref.bytes_to_ints(cref.gen_scalars("fp", 4, k))
import os  # noqa: E402
from halo2_b200 import lib as L  # noqa: E402
ways = int(os.environ.get("ACCUM_WAYS", "-1"))
if ways >= 0:
    L.check(L.init().h2_test_set_accum_ways(ways | (int(os.environ.get("ACCUM_LOG", "0")) << 8)))
    print(f"accum ways = {ways}", flush=True)
if os.environ.get("FAST"):
    L.check(L.init().h2_test_set_fast_fixed(int(os.environ["FAST"])))
    print("fast-fixed setting", os.environ["FAST"], flush=True)
for c in [int(a) for a in sys.argv[2:]] or [-1, 0, 13, 15, 17]:      # -1: digit-multiples table (direct sum); 0: automatic window
    try:
        t0 = time.perf_counter()
        params = h2.Params("vesta", k, g[:n], g[:n], g[n:n + 1], u=g[n + 1:n + 2], window_bits=max(c, 0), direct=(c < 0))
        print(f"c={c:2d}: Params setup {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"c={c}: {e}", flush=True)
        continue
    for _ in range(4):
        params.commit(pp, h2.Blind(5))
        params.commit_many(polys, [h2.Blind(5)] * 4)
    t0 = time.perf_counter()
    for _ in range(20):
        params.commit(pp, h2.Blind(5))
    t1 = time.perf_counter()
    for _ in range(10):
        params.commit_many(polys, [h2.Blind(5)] * 4)
    t2 = time.perf_counter()
    params.ipa_rounds(pp, 7, 9, lambda j, a, b: ch[j], lr, lr)
    t3 = time.perf_counter()
    for _ in range(3):
        params.ipa_rounds(pp, 7, 9, lambda j, a, b: ch[j], lr, lr)
    t4 = time.perf_counter()
    print(f"c={c:2d}: commit {(t1 - t0) / 20 * 1e3:.3f} ms  commit_many(4) {(t2 - t1) / 10 * 1e3:.3f} ms  ipa {(t4 - t3) / 3 * 1e3:.2f} ms", flush=True)
    params.close()
