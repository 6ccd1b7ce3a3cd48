"""Ad-hoc: a few fixed-base MSMs over digit-multiples tables (Params(direct=True)) for an ncu launch list."""
import sys

sys.path.insert(0, ".")
import halo2_b200 as h2  # noqa: E402
from oracle import cref  # noqa: E402  (input generation only)

k = int(sys.argv[1]) if len(sys.argv) > 1 else 14
n = 1 << k
g = cref.gen_points("vesta", 1, n + 2)
pp = cref.gen_scalars("fp", 2, n)
polys = [cref.gen_scalars("fp", 10 + i, n) for i in range(4)]
from halo2_b200 import lib as L  # noqa: E402
L.check(L.init().h2_test_set_graphs(0))
params = h2.Params("vesta", k, g[:n], g[:n], g[n:n + 1], u=g[n + 1:n + 2], direct=True)
for _ in range(3):
    params.commit(pp, h2.Blind(5))
params.commit_many(polys, [h2.Blind(5)] * 4)
params.close()
