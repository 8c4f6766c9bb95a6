"""Golden OUTPUT vectors of the oracle for BASELINE config 1's inputs (tests/golden/mnist_100_per_label_14x14.npz,
seeded random MPS of link dimension 10): one-step quantities of the per-bond path at selected bonds of the rightward
walk with W unchanged -- cost, per-label cost, #correct (fixedL.cc:280-344), cost and |r| after the first CG pass
(349-445), and the bond tensor rebuilt from its truncated SVD (519-521).  They are robust quantities (no CG
amplification), so they pin the oracle against regressions on CPU and give the GPU path committed numbers to meet.

  python tests/golden/make_golden_steps.py          -> tests/golden/config1_one_step.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fixedl_oracle as O  # noqa: E402
from tnml_b200 import data as D        # noqa: E402

BONDS = [1, 2, 3, 50, 96, 97, 98, 99, 150, 193, 194, 195]


def compute():
    here = os.path.dirname(os.path.abspath(__file__))
    g = np.load(os.path.join(here, "mnist_100_per_label_14x14.npz"))
    feat = O.features(g["sum4"].astype(np.float64) / (4 * 255.0))
    labels = g["labels"]
    W = D.random_mps(196, 2, 10, seed=1)
    ts = O.TrainStates(feat, labels)
    ts.init([None if w is None else w.copy() for w in W])
    out = {"bonds": np.array(BONDS), "w_checksum": np.array([sum(float(np.sum(w * w)) for w in W if w is not None)])}
    for b in range(1, 196):
        ts.set_bond(b)
        if b in BONDS:
            B = O.form_bond(W[b], W[b + 1])
            C, CL, ncor = O.quadcost(B, ts, detail=True)
            _, costs, rn = O.cgrad(B, ts, 2)
            Wb, Wb1, m, te = O.svd_split(B, b, 1, ts.jc, 20, 10, 1e-10)
            out[f"C{b}"] = np.array([C])
            out[f"CL{b}"] = np.asarray(CL, np.float64)
            out[f"ncor{b}"] = np.array([ncor])
            out[f"cost1_{b}"] = np.array([costs[0]])
            out[f"rn1_{b}"] = np.array([rn[0]])
            out[f"m{b}"] = np.array([m])
            out[f"newB{b}"] = O.form_bond(Wb, Wb1)
        ts.shiftE(W, b, "Fromleft")
    return out


if __name__ == "__main__":
    out = compute()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config1_one_step.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
