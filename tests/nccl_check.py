#!/usr/bin/env python3
"""Run under torchrun on N >= 2 GPUs (not collected by pytest: it needs one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/nccl_check.py
Checks the NCCL path's RESULT on hardware against the CPU oracle:
  1. halo2_b200.parallel.best_multiexp_sharded (host arrays in, one point out, every rank the same);
  2. the device-resident form bench.py times: h2_msm_dev -> all_gather_into_tensor -> h2_point_sum_dev on one stream.
Prints one JSON line from rank 0; exit status 1 on any mismatch."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from halo2_b200 import lib as L, parallel
    from oracle import cref, pasta
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    lib = L.init(local)
    ok = True
    report = {"world": world}
    for curve in ("pallas", "vesta"):
        c = pasta.CURVES[curve]
        for n in (0, 1, world - 1, 4097, 1 << 16):
            kb = cref.gen_scalars(c.scalar, 900 + n, n)
            pb = cref.gen_points(curve, 901 + n, n)
            got = parallel.best_multiexp_sharded(kb, pb, curve)
            want = cref.best_multiexp(curve, kb, pb) if n else np.zeros(64, dtype=np.uint8)
            good = bool((cref.jac_to_affine(curve, got) == want).all())
            ok &= good
            report[f"sharded_{curve}_{n}"] = good
        # device-resident pipeline: each rank's shard on its GPU, Montgomery bases, canonical scalars
        n = 1 << 15
        kb = cref.gen_scalars(c.scalar, 77, n * world)
        pb = cref.gen_points(curve, 78, n * world)
        sc = torch.from_numpy(kb[rank * n:(rank + 1) * n].copy()).to(dev)
        bs = torch.from_numpy(pb[rank * n:(rank + 1) * n].copy()).to(dev)
        sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        cid = L.CURVE_ID[curve]
        L.check(lib.h2_dev_convert(L.FIELD_ID[L.BASE_FIELD[curve]], ctypes.c_void_p(bs.data_ptr()), ctypes.c_size_t(2 * n), 1, sp))
        out = torch.zeros(96, dtype=torch.uint8, device=dev)
        gathered = torch.zeros(96 * world, dtype=torch.uint8, device=dev)
        final = torch.zeros(96, dtype=torch.uint8, device=dev)
        L.check(lib.h2_msm_dev(cid, ctypes.c_void_p(sc.data_ptr()), L.REPR_CANONICAL, ctypes.c_void_p(bs.data_ptr()), ctypes.c_size_t(n), 0,
                               ctypes.c_void_p(out.data_ptr()), sp))
        dist.all_gather_into_tensor(gathered, out)
        L.check(lib.h2_point_sum_dev(cid, ctypes.c_void_p(gathered.data_ptr()), ctypes.c_size_t(world), ctypes.c_void_p(final.data_ptr()), sp))
        L.check(lib.h2_dev_convert(L.FIELD_ID[L.BASE_FIELD[curve]], ctypes.c_void_p(final.data_ptr()), ctypes.c_size_t(3), 0, sp))
        torch.cuda.synchronize()
        good = bool((cref.jac_to_affine(curve, final.cpu().numpy()) == cref.best_multiexp(curve, kb, pb)).all())
        ok &= good
        report[f"device_pipeline_{curve}"] = good
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    report["all_ranks_ok"] = bool(flag.item())
    if rank == 0:
        print(json.dumps(report), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if report["all_ranks_ok"] else 1)


if __name__ == "__main__":
    main()
