"""Multi-GPU MSM: one process per GPU, contiguous index shards, ONE exchange.

best_multiexp is a sum of independent partial sums: rank g owns pairs
[g*n/G, (g+1)*n/G), runs the full single-GPU Pippenger on them, and the G partial results
(96-byte Jacobian points) are all-gathered over NCCL/NVLink and added on every rank in rank order.
Elliptic-curve addition is not an NCCL reduction op, hence all-gather + G-term EC sum rather than
all-reduce.  Scalars and bases never cross GPUs.  NTT stays single-GPU (north star)."""
from __future__ import annotations

import ctypes
from typing import Callable, Optional, Tuple

import numpy as np

from . import lib as _l


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of rank; sizes differ by at most one, empty shards allowed."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _local_msm(curve: str, coeffs: np.ndarray, bases: np.ndarray) -> np.ndarray:
    from .arithmetic import best_multiexp
    return best_multiexp(coeffs, bases, curve)


def _point_sum(curve: str, parts: np.ndarray) -> np.ndarray:
    lib = _l.init()
    out = np.zeros(96, dtype=np.uint8)
    _l.check(lib.h2_point_sum(_l.CURVE_ID[curve], _l.ptr(parts), ctypes.c_size_t(parts.shape[0]), _l.REPR_CANONICAL, _l.ptr(out)))
    return out


def best_multiexp_sharded(coeffs, bases, curve: str = "vesta", group=None,
                          local_msm: Optional[Callable] = None, point_sum: Optional[Callable] = None) -> np.ndarray:
    """best_multiexp over the process group: every rank passes the FULL arrays (or at least its
    own shard's rows) and receives the same Jacobian result.

    `local_msm` / `point_sum` default to the GPU engine; the CPU (gloo) tests inject oracle
    functions to exercise the sharding + exchange logic without a GPU."""
    import torch
    import torch.distributed as dist
    c = _l.as_u8(coeffs, 32)
    b = _l.as_u8(bases, 64)
    assert c.shape[0] == b.shape[0], "best_multiexp: coeffs.len() != bases.len()"
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(c.shape[0], rank, world)
    part = (local_msm or _local_msm)(curve, c[lo:hi], b[lo:hi])
    if world == 1:
        return part
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.from_numpy(np.ascontiguousarray(part)).to(dev)
    gathered = [torch.zeros(96, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    parts = np.stack([g.cpu().numpy() for g in gathered])
    return (point_sum or _point_sum)(curve, parts)


# ---- single-process multi-GPU (include/halo2_b200.h: h2_multi_*) ----------------------------------------------------
def multi_init(ngpu: int) -> int:
    """Binds the engine to `ngpu` devices of this process (the primary one from lib.init() first); returns the count."""
    lib = _l.init()
    _l.check(lib.h2_multi_init(int(ngpu)))
    return int(lib.h2_multi_count())


def best_multiexp_multi_gpu(coeffs, bases, curve: str = "vesta") -> np.ndarray:
    """best_multiexp (arithmetic.rs:143-180) sharded over the devices of multi_init inside ONE process: contiguous shards,
    parallel uploads, per-device Pippenger, partial results peer-written to the primary device and added there."""
    c = _l.as_u8(coeffs, 32)
    b = _l.as_u8(bases, 64)
    assert c.shape[0] == b.shape[0], "best_multiexp: coeffs.len() != bases.len()"
    out = np.zeros(96, dtype=np.uint8)
    _l.check(_l.init().h2_msm_multi_gpu(_l.CURVE_ID[curve], _l.ptr(c), _l.ptr(b), ctypes.c_size_t(c.shape[0]), _l.REPR_CANONICAL, _l.ptr(out)))
    return out


class MultiGpuBases:
    """A base vector resident in shards on the devices of multi_init (BASELINE configs[4]: bases pre-resident)."""

    def __init__(self, bases, curve: str = "vesta"):
        b = _l.as_u8(bases, 64)
        self.curve, self.n = curve, b.shape[0]
        self._h = ctypes.c_uint64(0)
        _l.check(_l.init().h2_multi_bases_register(_l.CURVE_ID[curve], _l.ptr(b), ctypes.c_size_t(self.n), _l.REPR_CANONICAL, ctypes.byref(self._h)))

    def msm(self, coeffs) -> np.ndarray:
        c = _l.as_u8(coeffs, 32)
        assert c.shape[0] == self.n, "best_multiexp: coeffs.len() != bases.len()"
        out = np.zeros(96, dtype=np.uint8)
        _l.check(_l.init().h2_msm_multi_registered(self._h, _l.ptr(c), ctypes.c_size_t(self.n), _l.REPR_CANONICAL, _l.ptr(out)))
        return out

    def close(self) -> None:
        if self._h.value:
            _l.check(_l.load().h2_multi_bases_release(self._h))
            self._h = ctypes.c_uint64(0)
