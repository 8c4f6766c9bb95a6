"""Worker of tests/test_multi_gpu.py: one rank of a 2-rank run through the C-ABI (one process per GPU,
images sharded like ParallelDo bounds, NCCL id exchanged through a file)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fixedl_oracle as O      # noqa: E402  (shard bounds + form_bond only: test infrastructure)
from tests.helpers import make_problem     # noqa: E402
from tnml_b200 import capi                 # noqa: E402


def main():
    rank, world, work = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    feat, labels, W = make_problem(N=10, NT=2600, m0=8)
    a, b = O.shard_bounds(world, feat.shape[0])[rank]
    h = capi.Handle(rank)
    h.set_images(feat[a:b], labels[a:b].astype(np.int32), feat.shape[0], a)
    h.set_mps(W)
    uidf = os.path.join(work, "uid")
    if rank == 0:
        uid = capi.comm_get_unique_id()
        with open(uidf + ".tmp", "wb") as f:
            f.write(uid)
        os.rename(uidf + ".tmp", uidf)
    else:
        for _ in range(600):
            if os.path.exists(uidf):
                break
            time.sleep(0.1)
        uid = open(uidf, "rb").read()
    if world > 1:
        h.comm_init_rank(world, rank, uid)
    h.init_envs()
    out = {}
    for bond in (1, 2, 3, 4):
        h.set_bond(bond)
        if bond in (3, 4):          # class L (b=3) and class C (b=4: label site is 5)
            B = O.form_bond(W[bond], W[bond + 1])
            h.bond_load(B)
            c, cl, nc = h.quadcost(False, 1e-4)
            h.cgrad(1, 1e-4)
            out[f"C{bond}"], out[f"CL{bond}"], out[f"nc{bond}"] = c, cl, nc
            out[f"B{bond}"] = h.bond_store()
            h.bond_load(B)
            costs, rn = h.cgrad(3, 1e-4)
            out[f"costs{bond}"], out[f"rn{bond}"] = np.array(costs), np.array(rn)
        h.shift_env(bond, capi.FROMLEFT)
    bc = h.comm_broadcast([3.5 + rank, -1.0 * rank], root=0) if world > 1 else [3.5, 0.0]
    out["bcast"] = np.array(bc)
    np.savez(os.path.join(work, f"rank{rank}_of{world}.npz"), **out)
    h.close()


if __name__ == "__main__":
    main()
