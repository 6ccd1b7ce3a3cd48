"""Builds the test-only host emulation of the device headers (see README.md)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "halo2_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libh2_kernel_emul.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, f) for f in sorted(os.listdir(HERE)) if f.endswith(".cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-fPIC", "-shared", "-pthread", "-I", CSRC,
           "-I", os.path.join(ROOT, "include"), "-DH2_HOST_EMUL=1", "-o", OUT] + srcs
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
