"""Ad-hoc: wall time of the proof-shaped k-replay (tests/prover_replay.py) on the GPU arm only, plus the per-kind split."""
import sys
import time
sys.path.insert(0, ".")
import halo2_b200 as h2  # noqa: E402
from oracle import cref, pasta  # noqa: E402
from tests import prover_replay as R  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 14
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n = 1 << k
pts = cref.gen_points("vesta", 50, n + 2)
g, w, u = pts[:n], pts[n:n + 1], pts[n + 1:n + 2]
gl = h2.lagrange_generators("vesta", k, g)
inp = R.replay_inputs(cref, k, 14)
omega = pasta.omega_for_k("fp", k)
gpu = R.GpuArm(h2, k, g, gl, w, u)
for _ in range(3):
    R.run(gpu, inp, k, omega); gpu.free()
ts = []
for _ in range(reps):
    t0 = time.time(); R.run(gpu, inp, k, omega); ts.append((time.time() - t0) * 1e3); gpu.free()
print(f"k={k}: replay ms " + " ".join(f"{t:.2f}" for t in ts) + f"   min {min(ts):.2f}", flush=True)
stamps = {}
orig_ipa = gpu.ipa
def timed_ipa(*a):
    t0 = time.time(); r = orig_ipa(*a); stamps["ipa"] = (time.time() - t0) * 1e3; return r
gpu.ipa = timed_ipa
R.run(gpu, inp, k, omega); gpu.free()
print(f"ipa (unsynced entry) {stamps['ipa']:.2f} ms")
proof = R.run(gpu, inp, k, omega); gpu.free()
gv = R.GpuVerifierArm(h2, k, g, gl, w, u, params=gpu.params)
ok = R.verify(gv, proof, k, omega)
ts = []
for _ in range(reps):
    t0 = time.time(); ok = R.verify(gv, proof, k, omega) and ok; ts.append((time.time() - t0) * 1e3)
print(f"verify accepted={ok} ms " + " ".join(f"{t:.2f}" for t in ts) + f"   min {min(ts):.2f}", flush=True)
t0 = time.time(); cv = R.CpuVerifierArm(cref, pasta, k, g, gl, w, u, 128); okc = R.verify(cv, proof, k, omega)
print(f"cpu verify accepted={okc} hot {cv.hot_s * 1e3:.1f} ms wall {(time.time() - t0) * 1e3:.1f} ms", flush=True)
gpu.close()
