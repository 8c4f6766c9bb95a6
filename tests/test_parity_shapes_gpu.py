"""Parity at the shapes the benchmark times (VERDICT r01 "What's weak" #1): ONE step at a time,
teacher-forced -- before every step the CUDA path is given the oracle's bond tensor / site tensors,
so the chaotic amplification of the reference's CG cannot accumulate and every
quantity can be compared sharply:

  environments, P, cost                       rel 1e-11 .. 1e-12
  first CG step  a*p = (|r|^2/pAp) r          rel 1e-9   (gradient contraction + pAp pass)
  cost after the first CG step                rel 1e-10
  truncated SVD: m equal, truncerr 1e-8, U*S*V rel 1e-10, isometry 1e-12

Link dimension 120 with 8192 images reaches krgemm2_kernel<4,3> / <2,3> with 8 k-tiles and more
than one 16-row tile per warp, fat_kernel_t<*,4>, krgram2 with 4 x 2 (ragged) output tiles and the
240 x 240 / 2400 x 240 SVD; 119 / 121 / 300 reach the 8-byte cp.async paths, the generic label
kernel and the large-m fallbacks.  The oracle runs shard by shard (tests/helpers.ShardedOracle).
"""
import numpy as np
import pytest

from oracle import fixedl_oracle as O
from tests.helpers import ShardedOracle, copy_mps, make_problem, rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from tnml_b200 import capi as c
    c.load_library()
    return c


class Walker:
    """CUDA handle and sharded oracle walking a chain in lockstep (same W on both sides)."""

    def __init__(self, capi, N, NT, m0, seed=5, chunk=1024):
        self.capi = capi
        feat, labels, W = make_problem(N=N, NT=NT, m0=m0, seed=seed)
        self.N, self.NT, self.W = N, NT, W
        self.so = ShardedOracle(feat, labels, chunk)
        self.so.init(W)
        self.h = capi.Handle(0)
        self.h.set_images(feat, labels.astype(np.int32))
        self.h.set_mps(W)
        self.h.init_envs()
        self.pos = 1

    def goto(self, b):
        assert b >= self.pos
        while self.pos < b:
            self.so.set_bond(self.pos)
            self.so.shiftE(self.W, self.pos, "Fromleft")
            self.h.set_bond(self.pos)
            self.h.shift_env(self.pos, self.capi.FROMLEFT)
            self.pos += 1
        self.so.set_bond(b)
        self.h.set_bond(b)

    def close(self):
        self.h.close()


def _one_step_checks(wk, capi, b, Npass, maxm, minm, check_svd=True):
    h, so, W = wk.h, wk.so, wk.W
    NT = wk.NT
    wk.goto(b)
    if b > 1:
        assert rel(h.get_env(b - 1), so.slot(b - 1)) < 1e-12          # left env (krgemm S=2)
    if b + 2 <= wk.N:
        assert rel(h.get_env(b + 2), so.slot(b + 2)) < 1e-12          # right env from init
    B = O.form_bond(W[b], W[b + 1])
    h.bond_form()
    assert rel(h.bond_store(), B) < 1e-13
    # forward: P, cost per label, ncorrect
    C, CL, ncor = so.quadcost(B)
    c, cl, nc = h.quadcost(False)
    assert abs(c - C) < 1e-11 * C and rel(cl, CL) < 1e-11
    lab, P = h.predict(want_P=True)
    Pref = so.project(B)
    assert rel(P, Pref) < 1e-11
    flips = int(np.sum(lab != O.argmax_first(np.abs(Pref))))
    assert flips <= 2 and abs(nc - ncor) <= 2, (flips, nc, ncor)          # exact ties of |P_l| only
    # one CG step: gradient contraction + pAp pass
    Bo, costs_o, rn_o, steps_o = so.cgrad(B, Npass)
    h.bond_load(B)
    h.cgrad(1)
    assert rel(h.bond_store() - B, steps_o[0]) < 1e-9
    # Npass passes: the first cost is a one-step quantity, later ones feel the CG's amplification
    h.bond_load(B)
    costs, rn = h.cgrad(Npass)
    assert len(costs) == len(costs_o)
    assert abs(costs[0] - costs_o[0]) < 1e-10 * costs_o[0]
    assert abs(rn[0] - rn_o[0]) < 1e-7 * rn_o[0]
    assert rel(costs, costs_o) < 1e-6
    if not check_svd:
        return
    # truncated SVD of the ORACLE's optimised bond tensor, both sweep directions
    for ha in (1, 2):
        Wb, Wb1, m, te = O.svd_split(Bo, b, ha, so.jc, maxm, minm, 1e-10)
        h.bond_load(Bo)
        gm, gte = h.svd_split(capi.FROMLEFT if ha == 1 else capi.FROMRIGHT, 1e-10, maxm, minm)
        assert gm == m, (b, ha, gm, m)
        assert abs(gte - te) <= 1e-8 * te + 1e-22 * np.linalg.norm(Bo) ** 2
        gWb, gWb1 = h.get_site(b), h.get_site(b + 1)
        assert rel(O.form_bond(gWb, gWb1), O.form_bond(Wb, Wb1)) < 1e-10
        iso = gWb if ha == 1 else gWb1
        if ha == 1:
            U = (np.transpose(iso, (0, 1, 3, 2)) if iso.ndim == 4 else iso).reshape(-1, m)
            assert np.abs(U.T @ U - np.eye(m)).max() < 1e-12
        else:
            V = iso.reshape(m, -1)
            assert np.abs(V @ V.T - np.eye(m)).max() < 1e-12
        # cost of the truncated tensor from the ORACLE's factors (gauge teacher-forcing)
        h.set_site(b, Wb)
        h.set_site(b + 1, Wb1)
        c2, _, _ = h.quadcost(True)
        C2, _, _ = so.quadcost(O.form_bond(Wb, Wb1))
        assert abs(c2 - C2) < 1e-11 * C2
    # leave both sides with the rightward factors of the original W (the walk continues with W)
    h.set_site(b, W[b])
    h.set_site(b + 1, W[b + 1])


@pytest.fixture(scope="module")
def walk120(capi):
    wk = Walker(capi, N=20, NT=8192, m0=120)
    yield wk
    wk.close()


@pytest.mark.parametrize("b", [8, 9, 10, 11, 12])
def test_one_step_m120(capi, walk120, b):
    """class L (b=8), C (b=9,10: label on site 10), R (b=11,12) with ml = mr = 120, 8192 images."""
    assert walk120.W[b].shape[0] == 120 and walk120.W[b + 1].shape[2] == 120
    _one_step_checks(walk120, capi, b, Npass=3 if b in (9, 10) else 4, maxm=120, minm=60)


@pytest.mark.parametrize("m0", [119, 121])
def test_one_step_odd_m(capi, m0):
    """odd link dimensions: 8-byte cp.async paths; 121 -> generic label-environment kernel (MCH=0)."""
    wk = Walker(capi, N=20, NT=4096, m0=m0, seed=6)
    try:
        for b in (8, 11):
            _one_step_checks(wk, capi, b, Npass=2, maxm=m0, minm=m0 // 2, check_svd=(b == 8))
    finally:
        wk.close()


def test_one_step_m300(capi):
    """config-5 link dimension: class L, C and R bond at ml = mr = 300 with 2048 images (large-m
    paths of the projection and of the SVD: 600 x 600 bond matrix)."""
    wk = Walker(capi, N=24, NT=2048, m0=300, seed=7, chunk=256)
    try:
        assert wk.W[10].shape[0] == 300 and wk.W[11].shape[2] == 300
        _one_step_checks(wk, capi, 10, Npass=2, maxm=300, minm=300)
        _one_step_checks(wk, capi, 11, Npass=2, maxm=300, minm=300, check_svd=False)
        _one_step_checks(wk, capi, 13, Npass=2, maxm=300, minm=300, check_svd=False)
    finally:
        wk.close()


def test_bond_update_call_one_step_m120(capi, walk120):
    """The fused `tnml_bond_update` entry point at the benchmark shape: CG costs of the first pass, m,
    and the cost after the SVD against the oracle run from the same W (one bond, so the CG's
    amplification stays at the 1e-6 level)."""
    wk = walk120
    b = 12
    wk.goto(b)
    W = wk.W
    B = O.form_bond(W[b], W[b + 1])
    Bo, costs_o, _, _ = wk.so.cgrad(B, 4)
    Wb, Wb1, m, te = O.svd_split(Bo, b, 1, wk.so.jc, 120, 60, 1e-10)
    C2, _, _ = wk.so.quadcost(O.form_bond(Wb, Wb1))
    p = capi.BondParams(4, 0.0, 1e-10, 1e-10, 120, 60, 0)
    r = wk.h.bond_update(b, 1, p)
    assert abs(r.cg_cost[0] - costs_o[0]) < 1e-10 * costs_o[0]
    assert r.newm == m
    assert abs(r.cost - C2) < 1e-5 * C2
    # the environment the update advanced -- tcgen05 kernel on the int8 planes that were cut for the
    # projection of this bond (re-used, not re-cut) -- against the oracle's advance through the SAME new site
    Wd = list(W)
    Wd[b] = wk.h.get_site(b)
    wk.so.shiftE(Wd, b, "Fromleft")
    assert rel(wk.h.get_env(b), wk.so.slot(b)) < 1e-12
    # restore the walk state: W(b), W(b+1) and the env slot the update advanced
    wk.h.set_site(b, W[b])
    wk.h.set_site(b + 1, W[b + 1])
    wk.h.set_bond(b)
    wk.h.shift_env(b, capi.FROMLEFT)
    wk.so.shiftE(W, b, "Fromleft")
    wk.pos = b + 1


def test_golden_mnist_sweep_teacher_forced(capi):
    """BASELINE config 1 on real MNIST (committed 1000-image 14x14 subset, maxm=20): a WHOLE sweep
    (390 bond updates), teacher-forced per bond -- the CUDA path gets the oracle's W(b), W(b+1) before
    each step and the oracle's factors after the SVD -- so every bond is a sharp one-step check."""
    import copy
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mnist_100_per_label_14x14.npz"))
    feat = O.features(g["sum4"].astype(np.float64) / (4 * 255.0))
    labels = g["labels"]
    from tnml_b200 import data as D
    W = D.random_mps(196, 2, 10, seed=1)
    N, NT, jc = 196, 1000, 98
    ts = O.TrainStates(feat, labels)
    ts.init(copy_mps(W))
    h = capi.Handle(0)
    h.set_images(feat, labels.astype(np.int32))
    h.set_mps(W)
    h.init_envs()
    worst = dict(step=0.0, step_noise=0.0, step_over_noise=0.0, env=0.0, cost1=0.0, newB=0.0, cost=0.0)
    prng = np.random.default_rng(7)
    for (b, ha) in O.sweep_schedule(N):
        ts.set_bond(b)
        h.set_bond(b)
        B = O.form_bond(W[b], W[b + 1])
        h.bond_form()
        assert rel(h.bond_store(), B) < 1e-12, (b, ha)
        # one-step quantities of the CG
        # Yardstick for the step: the reference's step length a = |G|^2 / pAp(G) is ill-conditioned on
        # real data (bond 18 of this sweep: |B| ~ 7e4, |G| ~ 1e-5; one-ulp noise on B moves a by 1e-9).
        # B and W are teacher-forced, but the environments are the CUDA path's own chain of advances
        # and drift from the oracle's by accumulated rounding (1e-16 ... 1e-14 along the sweep).  So the
        # oracle's step is re-evaluated with its inputs perturbed by exactly that much (B: one ulp,
        # environments: the measured deviation), and with 4 ParallelDo shards; the CUDA path must stay
        # spread is recorded (printed at the end: the CUDA path sits at 20-150x the three-draw estimate,
        # 1e-9 .. 6e-8) and the step itself is only asserted to 1e-6 here.  The sharp 1e-9 step checks
        # are the synthetic-data tests above; cost, m, truncation and the truncated tensor below ARE
        # asserted sharply on every bond.
        def oracle_step(Bx, tsx, nshard):
            tsx.bounds = O.shard_bounds(nshard, NT)
            Gx, _ = O._grad(Bx, tsx, 0.0, False)
            pAp = sum(float(np.sum(O.project(Gx, tsx, slice(a0, a1)) ** 2)) for (a0, a1) in tsx.bounds)
            tsx.bounds = O.shard_bounds(1, NT)
            return float(np.sum(Gx * Gx)) / pAp * Gx
        step = oracle_step(B, ts, 1)
        dev = {}
        for j in (b - 1, b + 2):
            if 1 <= j <= N and ts.slot[j] is not None:
                dev[j] = max(rel(h.get_env(j), ts.slot[j]), 1.1e-16)
                worst["env"] = max(worst["env"], dev[j])
                assert dev[j] < 1e-11, (b, ha, j, dev[j])
        spread = [rel(oracle_step(B, ts, 4), step)]
        for _ in range(3):
            ts2 = copy.copy(ts)
            ts2.slot = list(ts.slot)
            for j, dj in dev.items():
                ts2.slot[j] = ts.slot[j] * (1.0 + dj * prng.standard_normal(ts.slot[j].shape))
            spread.append(rel(oracle_step(B * (1.0 + 1.1e-16 * prng.standard_normal(B.shape)), ts2, 1), step))
        noise = max(spread)
        h.cgrad(1)
        e = rel(h.bond_store() - B, step)
        worst["step"] = max(worst["step"], e)
        worst["step_noise"] = max(worst["step_noise"], noise)
        worst["step_over_noise"] = max(worst["step_over_noise"], e / noise)
        assert e < max(1e-6, 60 * noise), (b, ha, e, noise)
        Bo, costs_o, _ = O.cgrad(B, ts, 4)
        h.bond_load(B)
        costs, _ = h.cgrad(4)
        if costs_o:
            e = abs(costs[0] - costs_o[0]) / costs_o[0]
            worst["cost1"] = max(worst["cost1"], e)
            assert e < 1e-9, (b, ha, e)
        Wb, Wb1, m, te = O.svd_split(Bo, b, ha, jc, 20, 10, 1e-10)
        h.bond_load(Bo)
        gm, gte = h.svd_split(capi.FROMLEFT if ha == 1 else capi.FROMRIGHT, 1e-10, 20, 10)
        assert gm == m, (b, ha, gm, m)
        assert abs(gte - te) <= 1e-7 * te + 1e-22 * float(np.sum(Bo * Bo)), (b, ha, gte, te)
        newB = O.form_bond(Wb, Wb1)
        e = rel(O.form_bond(h.get_site(b), h.get_site(b + 1)), newB)
        worst["newB"] = max(worst["newB"], e)
        assert e < 1e-10, (b, ha, e)
        W[b], W[b + 1] = Wb, Wb1
        h.set_site(b, Wb)
        h.set_site(b + 1, Wb1)
        C, _, ncor = O.quadcost(newB, ts, detail=True)
        c, _, nc = h.quadcost(True)
        e = abs(c - C) / C
        worst["cost"] = max(worst["cost"], e)
        assert e < 1e-10 and abs(nc - ncor) <= 1, (b, ha, e, nc, ncor)
        d = "Fromleft" if ha == 1 else "Fromright"
        ts.shiftE(W, b, d)
        h.shift_env(b, capi.FROMLEFT if ha == 1 else capi.FROMRIGHT)
    print("teacher-forced sweep, worst relative deviations:", worst, "final cost", C / NT)
    assert C / NT < 0.6          # the sweep learned something (starts at ~1.0)
    h.close()


def test_tcgen05_projection_matches_dmma_and_oracle(capi, walk120):
    """The tcgen05 int8 (error-free splitting) projection kernel -- default from 1024 images and link
    dimension 48 on -- against the FP64 mma.sync kernel and the oracle on forward quantities, for 8, 7
    and 6 planes (2^-57, 2^-50, 2^-43 of row max x column max), on a class-R bond (thin = right env)."""
    wk = walk120
    b = wk.pos                      # wherever the walk stands (class R, ml = mr = 120 or the tail)
    wk.goto(b)
    W = wk.W
    B = O.form_bond(W[b], W[b + 1])
    Pref = wk.so.project(B)
    C, _, _ = wk.so.quadcost(B)
    h = wk.h
    h.bond_load(B)
    res = {}
    for variant, ns, tol in ((2, 8, 1e-12), (3, 8, 1e-12), (3, 7, 1e-10), (3, 6, 1e-8)):
        h.set_option("krgemm_variant", variant)
        h.set_option("oz_slices", ns)
        c, _, _ = h.quadcost(False)
        _, P = h.predict(want_P=True)
        res[(variant, ns)] = P
        assert rel(P, Pref) < tol, (variant, ns, rel(P, Pref))
        assert abs(c - C) < 10 * tol * C
    assert rel(res[(3, 8)], res[(2, 8)]) < 1e-12
    assert rel(res[(3, 6)], res[(2, 8)]) > 0.0          # the 6-plane result is a different computation
    h.set_option("krgemm_variant", -1)
    h.set_option("oz_slices", 8)


def test_tcgen05_env_advance_bit_exact(capi):
    """The tcgen05 int8 path against its operation-by-operation numpy model (oracle/ozaki.py::oz_kernel_model:
    round-to-nearest 7-bit planes, exact integer level sums, exact merge, two roundings): BIT-IDENTICAL output
    on (a) the advance of a label-carrying environment (10 NT rows, one image per 10 rows, S = 2) and (b) the
    advance of a thin environment, both with link dimension 120 (K padded to 128, ragged column tiles).
    The model itself agrees with the plain float64 contraction to 4e-16 of sum |a||b| (tests/test_oracle.py)."""
    from oracle import ozaki
    N, NT = 20, 1024
    feat, labels, W = make_problem(N=N, NT=NT, m0=120, seed=9)
    h = capi.Handle(0)
    h.set_images(feat, labels.astype(np.int32))
    h.set_mps(W)
    h.init_envs()
    # (a) right environments are built N..3 by init_envs: slot 10 (label site) -> slot 9 through site 9
    prev, new = h.get_env(10), h.get_env(9)
    assert prev.shape == (NT, 10, 120) and new.shape == (NT, 10, 120)
    sel = 40
    Bm = np.ascontiguousarray(W[9].transpose(2, 1, 0)).reshape(2 * W[9].shape[2], W[9].shape[0])   # [(k, s)][j]
    ref = ozaki.oz_kernel_model(prev[:sel].reshape(sel * 10, 120), Bm, feat[:sel, 8, :], None, 2, 8, div=10)
    got = new[:sel].reshape(sel * 10, 120)
    assert np.array_equal(got, ref), float(np.abs(got - ref).max() / np.abs(ref).max())
    # the same numbers against plain float64: the int8 route is not the less accurate one
    plain = np.einsum("nlk,jsk,ns->nlj", prev[:sel], W[9], feat[:sel, 8, :]).reshape(sel * 10, 120)
    assert rel(got, plain) < 1e-14
    # (b) thin left environment: walk to bond 9, slot 8 -> slot 9
    for b in range(1, 9):
        h.set_bond(b)
        h.shift_env(b, capi.FROMLEFT)
    prev = h.get_env(8)
    h.set_bond(9)
    h.shift_env(9, capi.FROMLEFT)
    new = h.get_env(9)
    assert prev.shape == (NT, 120) and new.shape == (NT, 120)
    sel = 300
    Bm = W[9].reshape(W[9].shape[0] * 2, W[9].shape[2])                                           # [(k, s)][j]
    ref = ozaki.oz_kernel_model(prev[:sel], Bm, feat[:sel, 8, :], None, 2, 8)
    assert np.array_equal(new[:sel], ref), float(np.abs(new[:sel] - ref).max() / np.abs(ref).max())
    h.close()


def test_tcgen05_projection_and_label_kernel_bit_exact(capi):
    """Forward pass of a class-L bond at link dimension 120, P = B * t.v (fixedL.cc:318): the projection on the
    tcgen05 kernel (S = 4: both feature pairs as output-side weights) followed by the label-environment kernel,
    against the bit-level models of the two kernels (oracle/ozaki.py): BIT-IDENTICAL P on the first 40 images."""
    from oracle import ozaki
    N, NT, b = 20, 1024, 8
    feat, labels, W = make_problem(N=N, NT=NT, m0=120, seed=9)
    h = capi.Handle(0)
    h.set_images(feat, labels.astype(np.int32))
    h.set_mps(W)
    h.init_envs()
    for k in range(1, b):
        h.set_bond(k)
        h.shift_env(k, capi.FROMLEFT)
    h.set_bond(b)
    B = O.form_bond(W[b], W[b + 1])                      # [alpha][s][t][beta] = the device's class-L layout
    h.bond_load(B)
    h.quadcost(False)
    _, P = h.predict(want_P=True)
    thin, fat = h.get_env(b - 1), h.get_env(b + 2)
    assert thin.shape == (NT, 120) and fat.shape == (NT, 10, 120)
    sel = 40
    Q = ozaki.oz_kernel_model(thin[:sel], B.reshape(120 * 4, 120), feat[:sel, b - 1, :], feat[:sel, b, :], 4, 8)
    Pm = ozaki.fat_forward_model(Q, fat[:sel])
    assert np.array_equal(P[:sel], Pm), float(np.abs(P[:sel] - Pm).max() / np.abs(Pm).max())
    # and the modelled numbers are the float64 contraction to rounding
    plain = np.einsum("na,astb,ns,nt,nlb->nl", thin[:sel], B, feat[:sel, b - 1, :], feat[:sel, b, :], fat[:sel])
    assert rel(Pm, plain) < 1e-13
    h.close()
