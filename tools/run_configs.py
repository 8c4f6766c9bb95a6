"""Round-1 measurements on the BASELINE.json configurations (run on the GPU box):

  config 1  MNIST (committed 1000-image 14x14 golden subset), maxm=20, Nsweep=2: cost curve of the
            CUDA path vs the float64 oracle (two summation orders = the oracle's own noise floor)
  config 2  synthetic 14x14, NT=10000, maxm=50, 2 sweeps: sweep-average bond-updates/s
  config 3  synthetic 14x14, NT=60000, maxm=120, 2 sweeps: sweep-average + saturated-bond rate
  m=300     one class-L bond update at ml=mr=300 against the oracle (config 5 shape, small NT)
  config 5  per-bond times at m=300 with one GPU's shard of images (131072) on a window of bonds

  python tools/run_configs.py [1] [2] [3] [300]      -> profiles/configs_r01.txt
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fixedl_oracle as O  # noqa: E402   (checker)
from tnml_b200 import capi, data, fixedl  # noqa: E402

TAG = os.environ.get("TNML_PROFILE_TAG", "r02")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", f"configs_{TAG}.txt"), "a")   # copied to profiles/ afterwards


def say(*a):
    line = " ".join(str(x) for x in a)
    print(line, flush=True)
    out.write(line + "\n")
    out.flush()


def copy_mps(W):
    return [None if w is None else w.copy() for w in W]


def config1():
    g = np.load(os.path.join(ROOT, "tests", "golden", "mnist_100_per_label_14x14.npz"))
    feat = O.features(g["sum4"].astype(np.float64) / (4 * 255.0))
    labels = g["labels"]
    W = data.random_mps(196, 2, 10, seed=1)
    ts = fixedl.TrainStates(feat, labels.astype(np.int32))
    ts.init(W)
    t0 = time.perf_counter()
    gpu = fixedl.mldmrg(ts, 2, 20, 10, 1e-10)
    ts.h.synchronize()
    tg = time.perf_counter() - t0
    refs = []
    for ns in (1, 4):
        o = O.TrainStates(feat, labels, ns)
        o.init(copy_mps(W))
        t0 = time.perf_counter()
        refs.append(O.mldmrg(copy_mps(W), o, 2, 20, 10, 1e-10))
        tc = time.perf_counter() - t0
    say("== config 1: MNIST 1000 images (100/label), 14x14, maxm=20 minm=10, Nsweep=2, Npass=4, seeded W (m=10)")
    say(f"   GPU {len(gpu)} bond updates in {tg:.2f} s = {len(gpu) / tg:.1f} bond-updates/s ; numpy oracle (structured, "
        f"1 thread) {len(gpu) / tc:.2f} bond-updates/s")
    say("   bond# sweep half b   m   cost(GPU)      cost(oracle)   |GPU-or|/or  |or4-or1|/or1  ncor GPU/or")
    worst_g = worst_o = 0.0
    for k in range(len(gpu)):
        a, b, c = gpu[k], refs[0][k], refs[1][k]
        eg = abs(a["cost"] - b["cost"]) / b["cost"]
        eo = abs(c["cost"] - b["cost"]) / b["cost"]
        worst_g, worst_o = max(worst_g, eg), max(worst_o, eo)
        if k % 26 == 0 or k == len(gpu) - 1:
            say(f"   {k:4d}  {a['sweep']}    {a['half']}  {a['b']:3d} {a['m']:3d}  {a['cost']:.10f}  {b['cost']:.10f}  "
                f"{eg:.2e}     {eo:.2e}      {a['ncor']}/{b['ncor']}")
    say(f"   worst relative cost deviation over 2 sweeps: GPU vs oracle {worst_g:.2e} ; oracle(4 shards) vs oracle(1 shard) "
        f"{worst_o:.2e}")
    say(f"   final: GPU cost {gpu[-1]['cost']:.6f} ncor {gpu[-1]['ncor']} ; oracle {refs[0][-1]['cost']:.6f} ncor "
        f"{refs[0][-1]['ncor']} ; oracle-4 {refs[1][-1]['cost']:.6f} ncor {refs[1][-1]['ncor']}")
    ts.h.close()


def sweeps(NT, maxm, nsweep, tag):
    pix, labels = data.synthetic_digits(NT, 14, seed=20260925)
    feat = data.phi(pix)
    W = data.random_mps(196, 2, 10, seed=3)
    ts = fixedl.TrainStates(feat, labels)
    ts.init(W, reserve_m=maxm)
    say(f"== {tag}: synthetic 14x14, NT={NT}, maxm={maxm} minm={max(10, maxm // 2)}, Npass=4, start m=10")
    p = capi.BondParams(4, 0.0, 1e-10, 1e-10, maxm, max(10, maxm // 2), 0)
    for sw in range(1, nsweep + 1):
        t0 = time.perf_counter()
        sat_t, sat_n = 0.0, 0
        last = None
        for b, ha in fixedl.sweepnext(196):
            t1 = time.perf_counter()
            r = ts.h.bond_update(b, ha, p)
            dt = time.perf_counter() - t1
            if r.origm == maxm and r.newm == maxm and not (97 <= b <= 98):
                sat_t += dt
                sat_n += 1
            last = r
        dt = time.perf_counter() - t0
        say(f"   sweep {sw}: 390 bond updates in {dt:.2f} s = {390 / dt:.1f} bond-updates/s (sweep average, wall clock incl. "
            f"class-C and edge bonds); saturated class-L/R bonds (m_l=m_r={maxm}): {sat_n} at "
            f"{(sat_n / sat_t) if sat_t else 0:.1f} bond-updates/s; cost/NT {last.cost / NT:.6f}, train acc "
            f"{last.ncorrect * 100.0 / NT:.2f}%")
    ts.h.close()


def m300():
    N, NT, m = 24, 300, 300
    pix, labels = data.synthetic_digits(NT, 14, seed=9)
    feat = data.phi(pix[:, 86:86 + N])
    W = data.random_mps(N, 2, m, seed=5)
    o = O.TrainStates(feat, labels.astype(np.int64))
    o.init(copy_mps(W))
    h = capi.Handle(0)
    h.set_images(feat, labels)
    h.set_mps(W)
    h.init_envs()
    b = 10
    for bb in range(1, b):
        o.set_bond(bb)
        o.shiftE(W, bb, "Fromleft")
        h.set_bond(bb)
        h.shift_env(bb, capi.FROMLEFT)
    o.set_bond(b)
    Wo = copy_mps(W)
    oB = O.form_bond(Wo[b], Wo[b + 1])
    B, costs, _ = O.cgrad(oB, o, 4)
    Wb, Wb1, mm, te = O.svd_split(B, b, 1, 12, m, m // 2, 1e-10)
    Co, _, nco = O.quadcost(O.form_bond(Wb, Wb1), o, detail=True)
    t0 = time.perf_counter()
    r = h.bond_update(b, 1, capi.BondParams(4, 0.0, 1e-10, 1e-10, m, m // 2, 0))
    dt = time.perf_counter() - t0
    say(f"== m=300 check (config 5 shape): bond {b} ml={oB.shape[0]} mr={oB.shape[3]} NT={NT}: newm GPU {r.newm} / oracle {mm}; "
        f"cost GPU {r.cost / NT:.10f} / oracle {Co / NT:.10f} (rel {abs(r.cost - Co) / Co:.1e}); truncerr {r.truncerr:.3e}/{te:.3e}; "
        f"svd sweeps {r.svd_sweeps}; wall {dt * 1e3:.1f} ms")
    h.close()


def config5(NT=131072, m=300):
    """BASELINE config 5 shape (synthetic, m = 300) on a window of bonds: a class-L, the two class-C and
    a class-R bond at m_l = m_r = 300 with the images of one GPU's shard (1e6 / 8 = 125k)."""
    N = 24
    pix, labels = data.synthetic_digits(NT, 14, seed=20260925)
    feat = data.phi(pix[:, 86:86 + N])
    W = data.random_mps(N, 2, m, seed=5)
    h = capi.Handle(0)
    h.set_images(feat, labels)
    h.set_mps(W)
    h.set_option("reserve_m", m)
    h.init_envs()
    for bb in range(1, 10):
        h.set_bond(bb)
        h.shift_env(bb, capi.FROMLEFT)
    p = capi.BondParams(4, 0.0, 1e-10, 1e-10, m, m, 0)
    say(f"== config 5 window: synthetic N={N} sites, NT={NT} images on one GPU, maxm=minm={m}, Npass=4 (jc={N // 2})")
    for b in (10, 13, 14):
        if b == 13:      # pass the two class-C bonds without optimising them (measured separately below)
            for bb in (11, 12):
                h.set_bond(bb)
                h.shift_env(bb, capi.FROMLEFT)
        h.set_timing(True)
        h.stats(reset=True)
        t0 = time.perf_counter()
        r = h.bond_update(b, 1, p)
        h.synchronize()
        dt = time.perf_counter() - t0
        st = h.stats(reset=True)
        h.set_timing(False)
        cls = "L" if b <= N // 2 - 2 else ("C" if b <= N // 2 else "R")
        say(f"   bond {b} class {cls} m {r.origm}->{r.newm}: {dt * 1e3:.1f} ms = {1 / dt:.2f} bond-updates/s "
            f"({NT / dt / 1e6:.2f} M images*bonds/s); phases ms: proj {st.ms_proj:.1f} grad {st.ms_grad:.1f} fat {st.ms_fat:.1f} "
            f"svd {st.ms_svd:.1f} ({r.svd_sweeps} sweeps) shift {st.ms_shift:.1f}; cost/NT {r.cost / NT:.6f}")
    h.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["1", "2", "3", "300"]
    say(f"# run_configs {' '.join(which)}  ({capi.load_library().tnml_version().decode()})")
    if "1" in which:
        config1()
    if "300" in which:
        m300()
    if "2" in which:
        sweeps(10000, 50, 2, "config 2 (2 of 10 sweeps)")
    if "3" in which:
        sweeps(60000, 120, 2, "config 3 (2 of 20 sweeps)")
    if "5" in which:
        config5()
