"""Ad-hoc: one device-resident 2^k MSM with the default settings (target of ncu captures)."""
import sys
sys.path.insert(0, ".")
from tools.quick_gpu import msm_time, lib, L  # noqa: E402
k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
if len(sys.argv) > 3:
    L.check(lib.h2_test_set_batched_affine(int(sys.argv[2]), int(sys.argv[3])))
msm_time(k, reps=1)
