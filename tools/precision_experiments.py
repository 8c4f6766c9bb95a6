"""Precision experiments behind DESIGN.md section 3 (run on CPU, needs /root/reference MNIST
or uses the committed golden subset).  Compares cost curves of the float64 oracle under
 (a) a different summation order (1 vs 4 ParallelDo shards),
 (b) fp32 STORAGE of the environments, float64 math,
 (c) fp32-level rounding of the projected outputs P and of B (what 3xTF32 MMAs give)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixedl_oracle as O  # noqa: E402

g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "mnist_100_per_label_14x14.npz"))
feat = O.features(g["sum4"].astype(np.float64) / (4 * 255.0))
labels = g["labels"]
N = 196


def run(nshard, mode=None, maxb=120):
    W = O.random_mps(N, 2, 10, seed=1)
    ts = O.TrainStates(feat, labels, nshard=nshard)
    proj0 = O.project
    if mode in ("env32", "p32"):
        adv = ts._advance
        ts._advance = lambda *a, **k: adv(*a, **k).astype(np.float32).astype(np.float64)
    if mode == "p32":
        def proj(B, ts_, sl=slice(None), literal=False):
            return proj0(B.astype(np.float32).astype(np.float64), ts_, sl, literal).astype(np.float32).astype(np.float64)
        O.project = proj
    ts.init(W)
    r = O.mldmrg(W, ts, 1, 20, 10, 1e-10, max_bonds=maxb)
    O.project = proj0
    return r


a, b, c, d = run(1), run(4), run(1, "env32"), run(1, "p32")
print("bond  cost(f64,1 shard)   rel.dev 4 shards   rel.dev env-fp32   rel.dev P,B-fp32   ncorrect (f64 / 4sh / env32 / p32)")
for i in range(0, len(a), 6):
    e = lambda x: abs(a[i]["cost"] - x[i]["cost"]) / a[i]["cost"]
    print(f"{a[i]['b']:4d}  {a[i]['cost']:.10f}   {e(b):.2e}          {e(c):.2e}          {e(d):.2e}   "
          f"{a[i]['ncor']} / {b[i]['ncor']} / {c[i]['ncor']} / {d[i]['ncor']}")
