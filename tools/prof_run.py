"""Small fixed workload for ncu captures: one NTT 2^20 (Fp), one MSM 2^20 (Pallas), one MSM 2^14+1 (Vesta)."""
import sys
sys.path.insert(0, ".")
from tools.quick_gpu import ntt_time, msm_time  # noqa: E402
what = sys.argv[1:] or ["ntt", "msm", "small"]
if "ntt" in what:
    ntt_time(20, reps=1)
if "msm" in what:
    msm_time(20, reps=1)
if "small" in what:
    msm_time(14, curve=1, reps=1, extra=1)
