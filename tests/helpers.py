"""Shared builders for the parity tests (test infrastructure)."""
import numpy as np

from oracle import fixedl_oracle as O
from tnml_b200 import data as D


def make_problem(N=8, NT=300, m0=3, seed=5, L=None):
    """Small chain cut out of synthetic 14x14 digits: returns (feat, labels, W)."""
    L = L or 14
    pix, labels = D.synthetic_digits(NT, L, seed=seed)
    # take N sites from the busy middle of the raster
    start = (L * L - N) // 2
    feat = O.features(pix[:, start:start + N])
    W = D.random_mps(N, 2, m0, seed=seed)
    return feat, labels.astype(np.int64), W


def copy_mps(W):
    return [None if w is None else w.copy() for w in W]


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = max(np.abs(b).max(), 1e-300)
    return float(np.abs(a - b).max() / den)


class ShardedOracle:
    """The oracle on a large image set, shard by shard (images are independent: costs, gradients and
    pAp are plain sums over ParallelDo shards, paralleldo.h:32-43 / fixedL.cc:385,402,421).  Keeps the
    numpy temporaries of `fixedl_oracle.TrainStates` (A_n(j)*W(j) per image) at a few hundred MB so
    that link dimension 120-300 with thousands of images stays cheap.  Only oracle primitives are
    used; `cgrad` follows `fixedl_oracle.cgrad` line for line."""

    def __init__(self, feat, labels, chunk=1024):
        self.NT = feat.shape[0]
        self.ts = [O.TrainStates(feat[a:a + chunk], labels[a:a + chunk]) for a in range(0, self.NT, chunk)]
        self.jc = self.ts[0].jc

    def init(self, W):
        for t in self.ts:
            t.init(W)

    def set_bond(self, b):
        for t in self.ts:
            t.set_bond(b)

    def shiftE(self, W, b, direction):
        for t in self.ts:
            t.shiftE(W, b, direction)

    def slot(self, j):
        return np.concatenate([t.slot[j] for t in self.ts], axis=0)

    def project(self, B):
        return np.concatenate([O.project(B, t) for t in self.ts], axis=0)

    def quadcost(self, B, lam=0.0):
        C, CL, nc = 0.0, np.zeros(10), 0
        for t in self.ts:
            c, cl, n = O.quadcost(B, t, 0.0, detail=True)
            C, CL, nc = C + c, CL + cl, nc + n
        return C + lam * float(np.sum(B * B)), CL, nc

    def grad(self, B, lam=0.0):
        G, C = np.zeros_like(B), 0.0
        for t in self.ts:
            g, c = O._grad(B, t, 0.0, False)
            G, C = G + g, C + c
        if lam != 0.0:
            G = G - lam * B
        return G, C

    def cgrad(self, B, Npass=4, lam=0.0, cconv=1e-10):
        B = B.copy()
        r, _ = self.grad(B, lam)
        p = r.copy()
        costs, rnorms, steps = [], [], []
        for ps in range(1, Npass + 1):
            pAp = sum(float(np.sum(O.project(p, t) ** 2)) for t in self.ts) + lam * float(np.sum(p * p))
            a = float(np.sum(r * r)) / pAp
            B = B + a * p
            steps.append(a * p)
            if ps == Npass:
                break
            nr, C = self.grad(B, lam)
            beta = float(np.sum(nr * nr)) / float(np.sum(r * r))
            r = nr
            costs.append((C + lam * float(np.sum(B * B))) / self.NT)
            rn = float(np.sqrt(np.sum(r * r)))
            rnorms.append(rn)
            if rn < cconv:
                break
            p = r + beta * p
        return B, costs, rnorms, steps
