"""Multi-GPU agreement check (run under torchrun on N GPUs):
N ranks, images sharded with ParallelDo bounds, NCCL all-reduce of the bond
gradient inside libtnml_b200.so -- versus one rank holding all images.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29611 tools/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnml_b200 import capi, data, fixedl  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, NT = 16, 3001
    pix, labels = data.synthetic_digits(NT, 14, seed=5)
    feat = data.phi(pix[:, 90:90 + N])
    W = data.random_mps(N, 2, 6, seed=5)
    a, b = fixedl.bounds(world, NT)[rank]
    h = capi.Handle(local)
    h.set_images(feat[a:b], labels[a:b], NT, a)
    h.set_mps(W)
    uid = torch.zeros(capi.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.tensor(list(capi.comm_get_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    h.comm_init_rank(world, rank, bytes(uid.cpu().tolist()))
    h.init_envs()
    p = capi.BondParams(4, 1e-4, 1e-10, 1e-10, 12, 6, 0)
    sched = list(fixedl.sweepnext(N))
    mine = [h.bond_update(bb, ha, p) for bb, ha in sched]
    ok = True
    if rank == 0:
        ref = capi.Handle(local)
        ref.set_images(feat, labels)
        ref.set_mps(W)
        ref.init_envs()
        worst, early = 0.0, 0.0
        mism = []
        for k, (bb, ha) in enumerate(sched):
            r = ref.bond_update(bb, ha, p)
            e = abs(r.cost - mine[k].cost) / r.cost
            worst = max(worst, e)
            if k < 4:
                early = max(early, e)
            if r.newm != mine[k].newm:
                # a singular value sitting at the cutoff can fall on either side once the
                # trajectories have drifted apart; in the first bonds it must not
                mism.append((k, r.newm, mine[k].newm))
                if k < 4 or abs(r.newm - mine[k].newm) > 2:
                    ok = False
            if abs(r.ncorrect - mine[k].ncorrect) > NT // 100:
                ok = False
        # The all-reduce only changes the summation order (1e-16).  The first bonds must
        # therefore agree to ~1e-10; later the algorithm itself amplifies that noise by up to
        # 1e9 per bond (DESIGN.md "Precision"), exactly as two oracle runs with different
        # ParallelDo shard counts do, so only a loose bound is meaningful there.
        ok = ok and early < 1e-9 and worst < 5e-2
        print(f"multi_gpu_check world={world}: rel cost deviation vs single rank: first 4 bonds {early:.2e}, "
              f"whole sweep {worst:.2e}; newm differs at {mism} -> {'OK' if ok else 'FAIL'}")
        ref.close()
    # every rank must hold the same MPS (SVD runs replicated on bit-identical all-reduced data)
    Wm = np.concatenate([h.get_site(j).ravel() for j in range(1, N + 1)])
    t = torch.from_numpy(Wm).cuda()
    tmax, tmin = t.clone(), t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    same = bool((tmax == tmin).all().item())
    if rank == 0:
        print("replicated MPS bit-identical across ranks:", same)
    h.close()
    dist.destroy_process_group()
    if rank == 0 and not (ok and same):
        sys.exit(1)


if __name__ == "__main__":
    main()
