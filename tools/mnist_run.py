"""BASELINE config 3 on REAL MNIST (run on the GPU box; the idx files travel in baseline/_ref/MNIST/,
git-ignored, staged from the reference's mllib/MNIST by `mkdir -p baseline/_ref/MNIST && cp
/root/reference/mllib/MNIST/*ubyte* baseline/_ref/MNIST/`):

  (a) full training set (60000 images, 14x14 block means, features [1, x/4] with the reference's double
      /255), maxm=120 minm=60 cutoff 1e-10 Npass=4, `nsweep` sweeps from a seeded random MPS (m=10):
      cost/NT and training accuracy after every sweep, test error of `fulltest` on the 10000 t10k images;
  (b) a 6000-image subset (600 per label), maxm=50: cost-vs-bond curve of the CUDA path next to TWO
      float64 summation orders of the oracle (1 and 4 ParallelDo shards) -- their spread is the
      reference algorithm's own reproducibility.

  python tools/mnist_run.py [nsweep=3]            -> gpurun_out/mnist_<tag>.txt (copy to profiles/)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tnml_b200 import capi, data, fixedl  # noqa: E402

TAG = os.environ.get("TNML_PROFILE_TAG", "r02")
MN = os.path.join(ROOT, "baseline", "_ref", "MNIST")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", f"mnist_{TAG}.txt"), "w")   # copied to profiles/ afterwards: only gpurun_out/ travels back


def say(*a):
    line = " ".join(str(x) for x in a)
    print(line, flush=True)
    out.write(line + "\n")
    out.flush()


def load(kind, per_label):
    pix, labels = data.readMNIST(MN, kind, per_label)        # first per_label images of every label, /255 (mnist.h:495)
    return data.phi(data.reduce(pix, 14)), labels              # phi divides by 255 again (fixedL.cc:637-642)


def main():
    nsweep = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    if not os.path.isdir(MN):
        say(f"no MNIST files under {MN}: nothing to run")
        return
    say(f"# mnist_run ({capi.load_library().tnml_version().decode()})")
    # ---- (a) config 3 on the full training set
    feat, labels = load("Train", 7000)
    tfeat, tlabels = load("Test", 2000)
    NT = feat.shape[0]
    say(f"== config 3, real MNIST: {NT} training images (all), 14x14, maxm=120 minm=60 cutoff=1e-10 Npass=4; "
        f"{tfeat.shape[0]} test images; start: seeded random MPS m=10")
    W = data.random_mps(196, 2, 10, seed=1)
    ts = fixedl.TrainStates(feat, labels.astype(np.int32))
    ts.init(W, reserve_m=120)
    p = capi.BondParams(4, 0.0, 1e-10, 1e-10, 120, 60, 0)
    for sw in range(1, nsweep + 1):
        t0 = time.perf_counter()
        curve = []
        for b, ha in fixedl.sweepnext(196):
            r = ts.h.bond_update(b, ha, p)
            curve.append(r.cost / NT)
        dt = time.perf_counter() - t0
        Wt = ts.h.get_mps()
        ncor, _ = fixedl.fullTest(Wt, tfeat, tlabels.astype(np.int32), log=lambda s: None)
        say(f"   sweep {sw}: {dt:.2f} s = {390 / dt:.1f} bond-updates/s; cost/NT {r.cost / NT:.6f}; training accuracy "
            f"{100.0 * r.ncorrect / NT:.2f} %; TEST error {100.0 * (1 - ncor / tfeat.shape[0]):.2f} % ({tfeat.shape[0] - ncor} of "
            f"{tfeat.shape[0]} wrong); cost at bonds 1/98/195/390: {curve[0]:.5f} {curve[97]:.5f} {curve[194]:.5f} {curve[-1]:.5f}")
    ts.h.close()
    # ---- (b) 6000-image subset: CUDA path next to two oracle summation orders
    from oracle import fixedl_oracle as O          # checker only
    feat6, labels6 = load("Train", 600)
    NT6 = feat6.shape[0]
    W = data.random_mps(196, 2, 10, seed=1)
    ts = fixedl.TrainStates(feat6, labels6.astype(np.int32))
    ts.init(W, reserve_m=50)
    p = capi.BondParams(4, 0.0, 1e-10, 1e-10, 50, 25, 0)
    gpu = []
    for b, ha in fixedl.sweepnext(196):
        r = ts.h.bond_update(b, ha, p)
        gpu.append((r.cost / NT6, r.newm, int(r.ncorrect)))
    ts.h.close()
    refs = []
    for ns in (1, 4):
        o = O.TrainStates(feat6, labels6.astype(np.int64), ns)
        Wc = [None if w is None else w.copy() for w in W]
        o.init(Wc)
        t0 = time.perf_counter()
        refs.append(O.mldmrg(Wc, o, 1, 50, 25, 1e-10))
        say(f"   oracle ({ns} shard{'s' if ns > 1 else ''}): one sweep in {time.perf_counter() - t0:.1f} s")
    say(f"== {NT6}-image subset (600 per label), maxm=50 minm=25, one sweep: cost/NT after every 26th bond update")
    say("   bond#   m(GPU/or1/or4)  cost GPU        cost oracle-1   cost oracle-4   |GPU-or1|/or1  |or4-or1|/or1   ncor GPU/or1/or4")
    wg = wo = 0.0
    for k in range(len(gpu)):
        a, b1, b4 = gpu[k], refs[0][k], refs[1][k]
        eg = abs(a[0] - b1["cost"]) / b1["cost"]
        eo = abs(b4["cost"] - b1["cost"]) / b1["cost"]
        wg, wo = max(wg, eg), max(wo, eo)
        if k % 26 == 0 or k == len(gpu) - 1:
            say(f"   {k:4d}   {a[1]:3d}/{b1['m']:3d}/{b4['m']:3d}   {a[0]:.10f}  {b1['cost']:.10f}  {b4['cost']:.10f}  "
                f"{eg:.2e}       {eo:.2e}       {a[2]}/{b1['ncor']}/{b4['ncor']}")
    say(f"   worst relative cost deviation over the sweep: CUDA path vs oracle {wg:.2e}; oracle (4 shards) vs oracle (1 shard) {wo:.2e}")


if __name__ == "__main__":
    main()
