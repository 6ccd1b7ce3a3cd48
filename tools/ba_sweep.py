"""Ad-hoc: device-resident one-shot MSM time by batched-affine rounds / pairs per thread / kernel variant (msm.cuh K4a)."""
import sys
sys.path.insert(0, ".")
from tools.quick_gpu import msm_time, lib, L  # noqa: E402

ks = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [20]
for k in ks:
    L.check(lib.h2_test_set_batched_affine(0, 64))
    print("rounds=0: ", end=""); msm_time(k, reps=8)
    for variant in (0, 1, 2, 3):
        for rounds in (1, 2, 3):
            for target in (32, 64, 32 | 0x10000, 64 | 0x10000):
                if rounds == 1 and target & 0x10000:
                    continue
                L.check(lib.h2_test_set_batched_affine(rounds | ((variant + 1) << 8), target))
                print(f"variant={variant} rounds={rounds} target={target & 0xffff:3d}{'c' if target >> 16 else ' '}: ", end="")
                msm_time(k, reps=6)
    L.check(lib.h2_test_set_batched_affine(0, 32))
