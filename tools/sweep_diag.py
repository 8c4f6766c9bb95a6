"""Per-bond diagnostics of full sweeps (config 2/3 shapes): link dims, SVD sweeps, wall time per
bond and the CUDA-event phase breakdown per sweep.  Run on the GPU box:
  python tools/sweep_diag.py [NT] [maxm] [nsweep] [env_budget_gb]        -> gpurun_out/sweep_diag.txt"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tnml_b200 import capi, data, fixedl  # noqa: E402

NT = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
maxm = int(sys.argv[2]) if len(sys.argv) > 2 else 120
nsweep = int(sys.argv[3]) if len(sys.argv) > 3 else 2
budget_gb = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0     # environment tier: HBM budget for env slots
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "sweep_diag.txt"), "a")


def say(*a):
    line = " ".join(str(x) for x in a)
    print(line, flush=True)
    out.write(line + "\n")
    out.flush()


pix, labels = data.synthetic_digits(NT, 14, seed=20260925)
feat = data.phi(pix)
W = data.random_mps(196, 2, 10, seed=3)
ts = fixedl.TrainStates(feat, labels)
if budget_gb:
    ts.h.set_option("env_budget_gb", budget_gb)
ts.init(W, reserve_m=maxm)
h = ts.h
p = capi.BondParams(4, 0.0, 1e-10, 1e-10, maxm, max(10, maxm // 2), 0)
say(f"# sweep_diag NT={NT} maxm={maxm} env_budget_gb={budget_gb} ({capi.load_library().tnml_version().decode()})")
for sw in range(1, nsweep + 1):
    h.set_timing(True)
    h.stats(reset=True)
    rows = []
    t0 = time.perf_counter()
    for b, ha in fixedl.sweepnext(196):
        t1 = time.perf_counter()
        r = h.bond_update(b, ha, p)
        rows.append((b, ha, r.origm, r.newm, r.svd_sweeps, (time.perf_counter() - t1) * 1e3, r.npass_done))
    dt = time.perf_counter() - t0
    st = h.stats(reset=True)
    h.set_timing(False)
    a = np.array(rows)
    say(f"sweep {sw}: {len(rows)} bonds in {dt:.2f} s = {len(rows) / dt:.1f}/s; phases ms/bond: proj {st.ms_proj / 390:.2f} "
        f"grad {st.ms_grad / 390:.2f} fat {st.ms_fat / 390:.2f} svd {st.ms_svd / 390:.2f} shift {st.ms_shift / 390:.2f} "
        f"other {st.ms_other / 390:.2f}; launches {st.launches}; tier: {st.tier_evictions} evictions, {st.tier_fetches} fetches, "
        f"{st.tier_bytes / 1e9:.1f} GB over PCIe")
    say(f"   newm: min {int(a[:, 3].min())} median {int(np.median(a[:, 3]))} max {int(a[:, 3].max())}; svd sweeps: "
        f"median {int(np.median(a[:, 4]))} max {int(a[:, 4].max())} hist {np.bincount(a[:, 4].astype(int)).tolist()}")
    for k in list(range(0, 390, 15)) + [96, 97, 98, 99, 291, 292, 293]:
        b, ha, om, nm, ss, ms, npd = rows[k]
        say(f"   #{k:3d} b={b:3d} ha={ha} m {om:3d}->{nm:3d} svd_sweeps {ss:2d} npass {npd} wall {ms:7.2f} ms")
    sat = a[(a[:, 2] == maxm) & (a[:, 3] == maxm) & ((a[:, 0] < 97) | (a[:, 0] > 98))]
    if len(sat):
        say(f"   saturated class-L/R bonds: {len(sat)} at {1e3 / sat[:, 5].mean():.1f} bond-updates/s")
h.close()
