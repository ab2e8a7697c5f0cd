"""Real-NCCL parity check of the sharded drivers (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/check_multigpu.py

Every rank runs its slab with the CUDA engine + NCCL halo exchange and compares its owned rows,
bit for bit, with the single-domain CPU oracle result.  Prints one JSON line on rank 0.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (checker only)
from npbench_b200 import distributed as D  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    eng = D.B200Engine(lr)
    oracle.set_threads(max(1, oracle.max_threads() // world))
    rng = np.random.default_rng(7)
    ok = {}

    def dev(full, slab):
        return torch.from_numpy(np.ascontiguousarray(full[slab.row0:slab.row0 + slab.nloc])).cuda()

    ni, nj, ts = 64 * world + 37, 1000, 13
    A, B = rng.random((ni, nj)), rng.random((ni, nj))
    slab = D.Slab(ni, world, rank, D.JACOBI_MAX_BLOCK)
    lA, lB = dev(A, slab), dev(B, slab)
    D.jacobi_2d_sharded(eng, slab, ts, lA, lB)
    eng.synchronize()
    oracle.jacobi_2d(ts, A, B)
    ok["jacobi_2d"] = bool(np.array_equal(slab.owned(lA).cpu().numpy(), A[slab.lo:slab.hi]) and
                           np.array_equal(slab.owned(lB).cpu().numpy(), B[slab.lo:slab.hi]))

    # the same grid with the block passes served by the marching kernel (what big slabs get by default)
    A, B = rng.random((ni, nj)), rng.random((ni, nj))
    lA, lB = dev(A, slab), dev(B, slab)
    eng.lib.jacobi2d_set_mode(3)
    try:
        D.jacobi_2d_sharded(eng, slab, ts, lA, lB)
        eng.synchronize()
    finally:
        eng.lib.jacobi2d_set_mode(0)
    oracle.jacobi_2d(ts, A, B)
    ok["jacobi_2d_march"] = bool(np.array_equal(slab.owned(lA).cpu().numpy(), A[slab.lo:slab.hi]) and
                                 np.array_equal(slab.owned(lB).cpu().numpy(), B[slab.lo:slab.hi]))

    for name, march, H in (("heat_3d", False, 4), ("heat_3d_march", True, D.HEAT_GHOST)):
        shape, ts = (24 * world + 5, 40, 50), 8
        A, B = rng.random(shape), rng.random(shape)
        slab = D.Slab(shape[0], world, rank, H)
        lA, lB = dev(A, slab), dev(B, slab)
        D.heat_3d_sharded(eng, slab, ts, lA, lB, march=march)
        eng.synchronize()
        oracle.heat_3d(ts, A, B)
        ok[name] = bool(np.array_equal(slab.owned(lA).cpu().numpy(), A[slab.lo:slab.hi]) and
                        np.array_equal(slab.owned(lB).cpu().numpy(), B[slab.lo:slab.hi]))

    for name, march, H in (("fdtd_2d", False, 4), ("fdtd_2d_march", True, D.FDTD_GHOST)):
        nx, ny, tm = 50 * world + 11, 700, 13
        f = [rng.random((nx, ny)) for _ in range(3)]
        fict = rng.random(tm)
        slab = D.Slab(nx, world, rank, H)
        l = [dev(x, slab) for x in f]
        D.fdtd_2d_sharded(eng, slab, tm, l[0], l[1], l[2], fict, march=march)
        eng.synchronize()
        oracle.fdtd_2d(tm, f[0], f[1], f[2], fict)
        ok[name] = all(bool(np.array_equal(slab.owned(g).cpu().numpy(), w[slab.lo:slab.hi])) for g, w in zip(l, f))

    flags = torch.tensor([int(v) for v in ok.values()], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"check_multigpu": dict(zip(ok.keys(), [bool(x) for x in flags.tolist()])), "world": world}),
              flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if bool(flags.min().item()) else 1)


if __name__ == "__main__":
    main()
