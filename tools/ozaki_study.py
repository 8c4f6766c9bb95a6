"""CPU study for the round-2 route of DESIGN.md section 3/9: the float64 Khatri-Rao contractions of the
bond update emulated by an Ozaki-style error-free splitting into signed 7-bit slices with EXACT
integer products (what tcgen05 kind::i8 MMAs with int32 accumulators would compute), to find how
many slices the reference's conjugate-gradient update needs before its cost curve is
indistinguishable from the float64 one.

  A (rows scaled by 2^e_row)  ~  sum_i 2^(-7(i+1)) A_i ,   B (columns scaled)  ~  sum_j 2^(-7(j+1)) B_j
  A @ B  ~  sum_{i+j < s} 2^(-7(i+j+2)) (A_i @ B_j)         -> s(s+1)/2 integer GEMMs

Only the big contractions go through the emulation (projection: env x bond tensor, gradient:
env^T x back-propagated Z); the contraction with the label-carrying environment, the CG vector
algebra and the SVD stay float64, as they would on the device.  Yardstick = the oracle's own
spread between two summation orders (1 vs 4 ParallelDo shards).

  python tools/ozaki_study.py [max_bonds]        -> table on stdout (profiles/ozaki_study_r01.txt)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixedl_oracle as O  # noqa: E402

g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "mnist_100_per_label_14x14.npz"))
feat = O.features(g["sum4"].astype(np.float64) / (4 * 255.0))
labels = g["labels"]
N = 196
NSLICE = None      # None = plain float64


from oracle.ozaki import ozaki_matmul as _ozaki  # noqa: E402


def ozaki_matmul(A, B, s):
    return A @ B if s is None else _ozaki(A, B, s)


proj0, back0 = O.project, O.backproject


def project(B, ts, sl=slice(None), literal=False):
    if NSLICE is None or literal:
        return proj0(B, ts, sl, literal)
    l, x, y, r = O._lr(ts, sl)
    if B.ndim == 5:                       # class C
        ml, _, _, mr, nl = B.shape
        T = sum(x[:, s, None] * ozaki_matmul(l, B[:, s].reshape(ml, -1), NSLICE) for s in range(2))
        return np.einsum("ntbl,nt,nb->nl", T.reshape(-1, 2, mr, nl), y, r)
    if r.ndim == 3:                       # class L
        Q = sum((x[:, s] * y[:, t])[:, None] * ozaki_matmul(l, B[:, s, t, :], NSLICE) for s in range(2) for t in range(2))
        return np.einsum("nb,nlb->nl", Q, r)
    Q = sum((x[:, s] * y[:, t])[:, None] * ozaki_matmul(r, B[:, s, t, :].T, NSLICE) for s in range(2) for t in range(2))
    return np.einsum("na,nla->nl", Q, l)


def backproject(dP, Bshape, ts, sl=slice(None), literal=False):
    if NSLICE is None or literal:
        return back0(dP, Bshape, ts, sl, literal)
    l, x, y, r = O._lr(ts, sl)
    G = np.zeros(Bshape)
    if len(Bshape) == 5:
        nl = Bshape[4]
        Zf = np.einsum("nl,nt,nb->ntbl", dP, y, r).reshape(dP.shape[0], -1)
        for s in range(2):
            G[:, s] = ozaki_matmul((l * x[:, s, None]).T, Zf, NSLICE).reshape(Bshape[0], 2, Bshape[3], nl)
        return G
    if r.ndim == 3:
        Z = np.einsum("nl,nlb->nb", dP, r)
        for s in range(2):
            for t in range(2):
                G[:, s, t, :] = ozaki_matmul((l * (x[:, s] * y[:, t])[:, None]).T, Z, NSLICE)
        return G
    Z = np.einsum("nl,nla->na", dP, l)
    for s in range(2):
        for t in range(2):
            G[:, s, t, :] = ozaki_matmul((Z * (x[:, s] * y[:, t])[:, None]).T, r, NSLICE)
    return G


def run(nshard, nslice, maxb):
    global NSLICE
    NSLICE = nslice
    O.project, O.backproject = project, backproject
    W = O.random_mps(N, 2, 10, seed=1)
    ts = O.TrainStates(feat, labels, nshard=nshard)
    ts.init(W)
    r = O.mldmrg(W, ts, 1, 20, 10, 1e-10, max_bonds=maxb)
    O.project, O.backproject = proj0, back0
    NSLICE = None
    return r


if __name__ == "__main__":
    maxb = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    # single-GEMM accuracy on an environment-like operand
    rng = np.random.default_rng(0)
    A = rng.standard_normal((512, 120)) * np.exp(rng.standard_normal((512, 1)) * 3)
    B = rng.standard_normal((120, 120))
    ref = A @ B
    print("slices  integer GEMMs  max |err| / (|A| |B|) of one 512x120x120 product")
    for s in (3, 4, 5, 6, 7, 8):
        err = np.max(np.abs(ozaki_matmul(A, B, s) - ref) / (np.abs(A) @ np.abs(B)))
        print(f"  {s}        {s * (s + 1) // 2:3d}          {err:.2e}")
    base, reorder = run(1, None, maxb), run(4, None, maxb)
    res = {s: run(1, s, maxb) for s in (4, 5, 6, 7, 8)}
    print("\nbond  cost(f64)      rel.dev 4 shards " + " ".join(f"  {s} slices " for s in res) + "   ncorrect f64 / 4sh / " +
          " / ".join(str(s) for s in res))
    worst = {k: 0.0 for k in ["re"] + list(res)}
    for i in range(len(base)):
        e = lambda x: abs(base[i]["cost"] - x[i]["cost"]) / base[i]["cost"]
        worst["re"] = max(worst["re"], e(reorder))
        for s in res:
            worst[s] = max(worst[s], e(res[s]))
        if i % 6 == 0 or i == len(base) - 1:
            print(f"{base[i]['b']:4d}  {base[i]['cost']:.10f}   {e(reorder):.2e}       " +
                  "   ".join(f"{e(res[s]):.2e}" for s in res) + f"   {base[i]['ncor']} / {reorder[i]['ncor']} / " +
                  " / ".join(str(res[s][i]["ncor"]) for s in res))
    print("\nworst relative cost deviation over the run: 4 shards (yardstick) %.2e ; " % worst["re"] +
          " ; ".join(f"{s} slices {worst[s]:.2e}" for s in res))
