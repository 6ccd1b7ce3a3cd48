"""Ad-hoc: wall time of the device IPA round loop at k (default 14) + one commit for scale."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import halo2_b200 as h2  # noqa: E402
from oracle import cref  # noqa: E402  (input generation only)

k = int(sys.argv[1]) if len(sys.argv) > 1 else 14
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = 1 << k
g = cref.gen_points("vesta", 1, n + 2)
params = h2.Params("vesta", k, g[:n], g[:n], g[n:n + 1], u=g[n + 1:n + 2])
pp = cref.gen_scalars("fp", 2, n)
ch = cref.bytes_to_ints(cref.gen_scalars("fp", 3, k))
lr = cref.bytes_to_ints(cref.gen_scalars("fp", 4, k))
for it in range(reps):
    t0 = time.time()
    params.commit(pp, h2.Blind(5))
    t1 = time.time()
    stamps = []
    params.ipa_rounds(pp, 7, 9, lambda j, a, b: (stamps.append(time.time()), ch[j])[1], lr, lr)
    t2 = time.time()
    print(f"commit {1e3 * (t1 - t0):.3f} ms   ipa {1e3 * (t2 - t1):.3f} ms  rounds(ms): " +
          " ".join(f"{1e3 * (b - a):.2f}" for a, b in zip([t1] + stamps[:-1], stamps)), flush=True)
