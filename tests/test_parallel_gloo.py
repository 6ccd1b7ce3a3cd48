"""world_size-2 gloo test (CPU) of the N>1 MSM path: sharding + all-gather of 96-byte partials +
point sum, with the oracle standing in for the per-rank GPU MSM."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from halo2_b200.parallel import best_multiexp_sharded, shard_range
    from oracle import cref
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    curve = "pallas"
    kb = cref.gen_scalars("fq", 1234, n)
    pb = cref.gen_points(curve, 4321, n)

    def local_msm(cv, c, b):
        out = np.zeros(96, dtype=np.uint8)   # affine -> Jacobian with z = 1 (identity: z = 0)
        aff = cref.best_multiexp(cv, c, b, 2)
        if aff.any():
            out[:64] = aff
            out[64] = 1
        return out

    def point_sum(cv, parts):
        acc = np.zeros(64, dtype=np.uint8)
        for p in parts:
            acc = cref.point_add(cv, acc, cref.jac_to_affine(cv, p))
        out = np.zeros(96, dtype=np.uint8)
        if acc.any():
            out[:64] = acc
            out[64] = 1
        return out

    got = best_multiexp_sharded(kb, pb, curve, local_msm=local_msm, point_sum=point_sum)
    want = cref.best_multiexp(curve, kb, pb, 2)
    ok = bool((cref.jac_to_affine(curve, got) == want).all())
    lo, hi = shard_range(n, rank, world)
    ret[rank] = (ok, lo, hi)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [257, 1])
def test_sharded_msm_world2(n):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n, ret), nprocs=world, join=True)
    assert all(ret[r][0] for r in range(world))
    spans = [ret[r][1:] for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n and spans[0][1] == spans[1][0]


def test_shard_range_partition():
    from halo2_b200.parallel import shard_range
    for n in (0, 1, 7, 8, 1 << 20, (1 << 24) + 3):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
