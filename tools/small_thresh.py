"""Ad-hoc: one-shot MSM time around the lane-cooperative / thread-per-item threshold."""
import sys
sys.path.insert(0, ".")
from tools.quick_gpu import msm_time, lib, L  # noqa: E402
for lg in (20, 21, 22):
    L.check(lib.h2_test_set_accum_ways(0 | (lg << 8)))
    print(f"threshold 2^{lg}:")
    for k in (15, 16, 17, 18):
        msm_time(k, reps=10)
