"""Compiles the sm_100a CUDA library in-tree: halo2_b200/_lib/libhalo2_b200.so.

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box with
the gpurun snapshot.  There is no other backend and no CPU fallback."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIB_DIR, "libhalo2_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-split-compile", "0",
]
OBJ_DIR = os.path.join(LIB_DIR, "obj")


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def headers():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    out.append(os.path.join(os.path.dirname(HERE), "include", "halo2_b200.h"))
    return out


def _obj(src: str) -> str:
    return os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")


def _stale(target: str, deps) -> bool:
    return not os.path.exists(target) or any(os.path.getmtime(target) < os.path.getmtime(d) for d in deps)


def _deps_of(src: str):
    """The headers `src` includes, from the dependency file nvcc wrote next to its object (all headers when there is none)."""
    dfile = _obj(src) + ".d"
    if not os.path.exists(dfile):
        return [src] + headers()
    with open(dfile) as f:
        toks = f.read().replace("\\\n", " ").split()
    root = os.path.dirname(HERE)
    deps = [t for t in toks[1:] if t.startswith(root) and os.path.exists(t)]
    return deps or [src] + headers()


def up_to_date() -> bool:
    return not _stale(LIB, sources() + headers())


def build(force: bool = False, verbose: bool = False) -> str:
    """Every csrc/*.cu is one translation unit: they compile in parallel (one nvcc each) and only the stale ones are rebuilt
    -- all of them when a header changed -- then link into one shared library."""
    if not force and up_to_date():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libhalo2_b200.so (and there is no CPU fallback)")
    os.makedirs(OBJ_DIR, exist_ok=True)
    todo = [s for s in sources() if force or _stale(_obj(s), _deps_of(s))]
    procs = []
    for s in todo:
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-MD", "-MF", _obj(s) + ".d", "-c", "-o", _obj(s), s]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            failed.append(os.path.basename(s))
    if failed:
        raise RuntimeError("nvcc failed on " + ", ".join(failed))
    wanted = {_obj(s) for s in sources()}
    for f in os.listdir(OBJ_DIR):          # objects of sources that no longer exist
        if os.path.join(OBJ_DIR, f) not in wanted and os.path.join(OBJ_DIR, f[:-2]) not in wanted:
            os.remove(os.path.join(OBJ_DIR, f))
    tmp = LIB + ".tmp.so"   # built aside and renamed: a gpurun snapshot never sees a half-written library
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", tmp] + sorted(wanted)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("linking libhalo2_b200.so failed")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
