"""Ad-hoc: device-resident one-shot MSM time by size and window (input to msm_default_window)."""
import sys
sys.path.insert(0, ".")
from tools.quick_gpu import msm_time  # noqa: E402

ks = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [8, 10, 12, 14, 16, 18]
cs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 6, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17]
for k in ks:
    for c in cs:
        if c and (c > k + 3 or c < k - 8):
            continue
        msm_time(k, c=c, reps=5)
