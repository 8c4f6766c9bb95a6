"""Host-side mirror of the reference's `fixedL` interface, over the C-ABI.

Same names and argument meaning as fixedL.cc: `TrainStates` (64-274: init,
setBond, shiftE), `cgrad` (349-445), `quadcost` (280-344), `mldmrg` (451-570),
`sweepnext`, and paralleldo.h's shard bounds.  All arithmetic happens in
libtnml_b200.so on the GPU; this file only sequences calls.
"""
from __future__ import annotations

import numpy as np

from . import capi

Fromleft, Fromright = capi.FROMLEFT, capi.FROMRIGHT


def bounds(nshard, ntask):
    """ParallelDo(Nthread, Ntask) (paralleldo.h:32-43)."""
    th = ntask // nshard
    b = [(n * th, (n + 1) * th) for n in range(nshard)]
    b[-1] = (b[-1][0], ntask)
    return b


def sweepnext(N):
    """b = 1..N-1 (ha=1) then N-1..1 (ha=2), like `sweepnext(b,ha,N)`."""
    for b in range(1, N):
        yield b, 1
    for b in range(N - 1, 0, -1):
        yield b, 2


class TrainStates:
    """Training images of one shard, resident on one GPU (fixedL.cc:64-274)."""

    def __init__(self, feat, labels, device=0, NT_global=0, first=0):
        self.h = capi.Handle(device)
        self.h.set_images(feat, labels, NT_global, first)
        self.N = self.h.N
        self.NT = NT_global or self.h.NT

    def size(self):
        return self.NT

    def init(self, W, reserve_m=0):
        """fixedL.cc:122-157.  reserve_m (normally maxm): size the environment slots for that link
        dimension at once so they never have to grow during the sweeps."""
        self.h.set_mps(W)
        if reserve_m:
            self.h.set_option("reserve_m", reserve_m)
        self.h.init_envs()

    def setBond(self, b):
        self.h.set_bond(b)

    def shiftE(self, W, b, direction):
        """fixedL.cc:192-233 (W lives on the device; the argument is kept for
        signature parity)."""
        self.h.shift_env(b, direction)


def quadcost(B, ts: TrainStates, lam=0.0):
    """fixedL.cc:280-344: returns (C un-normalised, per-label costs, ncorrect)."""
    if B is not None:
        ts.h.bond_load(B)
    return ts.h.quadcost(False, lam)


def cgrad(B, ts: TrainStates, Npass=4, lam=0.0, cconv=1e-10):
    """fixedL.cc:349-445: returns (B, costs per pass, |r| per pass)."""
    ts.h.bond_load(B)
    costs, rn = ts.h.cgrad(Npass, lam, cconv)
    return ts.h.bond_store(), costs, rn


def mldmrg(ts: TrainStates, Nsweep, maxm, minm, cutoff, Npass=4, lam=0.0, cconv=1e-10,
           do_rel_cutoff=False, log=None, max_bonds=None):
    """fixedL.cc:451-570.  The MPS stays on the device (ts.h.get_mps())."""
    p = capi.BondParams(Npass, lam, cconv, cutoff, maxm, minm, int(do_rel_cutoff))
    NT = ts.size()
    out = []
    for sw in range(1, Nsweep + 1):
        if log:
            log(f"\nSweep {sw} maxm={maxm} minm={minm}")
        for b, ha in sweepnext(ts.N):
            c = b if ha == 1 else b + 1
            r = ts.h.bond_update(b, ha, p)
            if log:
                log(f"Sweep {sw} Half {ha} Bond {c}")
                for k in range(min(r.npass_done, 8)):
                    log(f"  Cost = {r.cg_cost[k]:.10f}")
                log(f"SVD trunc err = {r.truncerr:.2E}")
                log(f"Original m={r.origm}, New m={r.newm}")
                log(f"Percent correct = {r.ncorrect * 100.0 / NT:.4f}%, # incorrect = {NT - r.ncorrect}/{NT}")
                log(f"--> After SVD, Cost = {r.cost / NT:.10f}")
            out.append(dict(sweep=sw, half=ha, b=b, c=c, cost=r.cost / NT, ncor=int(r.ncorrect),
                            m=r.newm, truncerr=r.truncerr, cg_costs=list(r.cg_cost[:r.npass_done]),
                            cg_rnorms=list(r.cg_rnorm[:r.npass_done]), dB=r.dB, Bnorm=r.normB,
                            svd_sweeps=r.svd_sweeps))
            if max_bonds is not None and len(out) >= max_bonds:
                return out
    return out


def fullTest(W, feat, labels, device=0, log=print):
    """util.h:123-200 `fullTest` / fulltest.cc: classify a test set with the MPS W and print the
    reference's report lines.  Returns (ncorrect, predictions)."""
    h = capi.Handle(device)
    h.set_images(feat, labels)
    h.set_mps(W)
    pred, ncor = h.fulltest()
    h.close()
    labels = np.asarray(labels)
    nte = len(labels)
    ninc = nte - ncor
    if log:
        log(f"{ncor}/{nte} correct ({ncor * 100.0 / nte:.2f}%), {ninc}/{nte} incorrect ({ninc * 100.0 / nte:.2f}%)")
        for l in range(10):
            nt = int(np.sum(labels == l))
            if nt == 0:
                continue
            ni = int(np.sum((labels == l) & (pred != l)))
            log(f"  Digit {l} {nt - ni}/{nt} correct ({(nt - ni) * 100.0 / nt:.2f}%), {ni}/{nt} incorrect ({ni * 100.0 / nt:.2f}%)")
        log(f"Total # test images = {nte}")
    return ncor, pred
