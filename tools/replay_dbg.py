"""Ad-hoc: why bench.py's replay number is higher than tools/replay_time.py's -- same run after (a) nothing, (b) torch CUDA init, (c) a 128-thread CPU burst."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import halo2_b200 as h2
from oracle import cref, pasta
from tests import prover_replay as R
k = 14; n = 1 << k
pts = cref.gen_points("vesta", 50, n + 2)
g, w, u = pts[:n], pts[n:n + 1], pts[n + 1:n + 2]
gl = h2.lagrange_generators("vesta", k, g)
inp = R.replay_inputs(cref, k, 14)
omega = pasta.omega_for_k("fp", k)
gpu = R.GpuArm(h2, k, g, gl, w, u)
def t(label, reps=5):
    for _ in range(2):
        R.run(gpu, inp, k, omega); gpu.free()
    ts = []
    for _ in range(reps):
        t0 = time.time(); R.run(gpu, inp, k, omega); ts.append((time.time() - t0) * 1e3); gpu.free()
    print(f"{label}: " + " ".join(f"{x:.2f}" for x in ts), flush=True)
t("plain")
import torch
x = torch.zeros(1 << 20, device="cuda"); torch.cuda.synchronize()
t("after torch cuda init")
kb = cref.gen_scalars("fq", 3, 1 << 18); pb = cref.gen_points("pallas", 4, 1 << 18)
cref.best_multiexp("pallas", kb, pb)
t("after a 128-thread CPU MSM")
big = [torch.empty(256 << 20, dtype=torch.uint8, device="cuda") for _ in range(8)]
t("with 2 GiB of torch allocations alive")
import gc; gc.disable()
t("gc disabled")
gpu.close()
