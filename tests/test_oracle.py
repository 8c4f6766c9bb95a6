"""CPU tests that pin the oracle (no GPU): two formulations agree, invariants
hold, the data path matches known facts of the MNIST files."""
import hashlib
import os

import numpy as np
import pytest

from oracle import fixedl_oracle as O
from tests.helpers import copy_mps, make_problem, rel


def test_mnist_known_answers(ref_mnist_dir):
    md5 = {"train-images-idx3-ubyte": "6bbc9ace898e44ae57da46a324031adb",
           "train-labels-idx1-ubyte": "a25bea736e30d166cdddb491f175f624"}
    for f, want in md5.items():
        got = hashlib.md5(open(os.path.join(ref_mnist_dir, f), "rb").read()).hexdigest()
        assert got == want
    labs = O.read_idx(os.path.join(ref_mnist_dir, "train-labels-idx1-ubyte"))
    assert np.bincount(labs, minlength=10).tolist() == [5923, 6742, 5958, 6131, 5842, 5421, 5918, 6265, 5851, 5949]
    data, labels, idx = O.read_mnist(ref_mnist_dir, "Train", 100)
    assert data.shape == (1000, 784) and idx[:10].tolist() == list(range(10)) and idx[-1] == 1105
    assert np.bincount(labels, minlength=10).tolist() == [100] * 10
    assert data.max() <= 1.0


def test_golden_mnist_subset_matches_reference(ref_mnist_dir):
    """tests/golden/mnist_100_per_label_14x14.npz was produced by
    tests/golden/make_golden.py from the reference's MNIST files."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mnist_100_per_label_14x14.npz"))
    data, labels, _ = O.read_mnist(ref_mnist_dir, "Train", 100)
    d14 = O.reduce_image(data, 14)
    assert np.array_equal(g["labels"], labels)
    assert np.allclose(g["sum4"].astype(np.float64) / (4 * 255.0), d14, rtol=0, atol=1e-15)


def test_reduce_and_phi():
    img = np.arange(16, dtype=np.float64).reshape(1, 16) / 255.0
    r = O.reduce_image(img, 2)
    # 4x4 -> 2x2 block means, raster order
    assert np.allclose(r * 255.0, [[2.5, 4.5, 10.5, 12.5]])
    f = O.phi(np.array([0.0, 1.0]))
    assert np.allclose(f, [[1.0, 0.0], [1.0, 1.0 / 255.0 / 4.0]])
    with pytest.raises(ValueError):
        O.phi(np.array([-1.0]))


def test_shard_bounds():
    assert O.shard_bounds(3, 10) == [(0, 3), (3, 6), (6, 10)]
    assert O.shard_bounds(1, 7) == [(0, 7)]


def test_sweep_schedule():
    s = O.sweep_schedule(5)
    assert s == [(1, 1), (2, 1), (3, 1), (4, 1), (4, 2), (3, 2), (2, 2), (1, 2)]


def test_truncation_rule():
    P = np.array([1.0, 0.5, 1e-3, 1e-12, 1e-13, 1e-14])
    assert O.truncate_spectrum(P, 10, 1, 1e-10)[0] == 3
    assert O.truncate_spectrum(P, 2, 1, 1e-10)[0] == 2            # maxm wins
    assert O.truncate_spectrum(P, 10, 5, 1e-10)[0] == 5           # minm wins
    m, te = O.truncate_spectrum(P, 2, 1, 1e-10)
    assert np.isclose(te, 1e-3 + 1e-12 + 1e-13 + 1e-14)
    m, te = O.truncate_spectrum(P, 10, 1, 1e-10, do_rel_cutoff=True)
    assert m == 3 and np.isclose(te, (1e-12 + 1e-13 + 1e-14) / P.sum())


def _walk(ts, W, upto):
    """advance the right sweep without optimising, so that bond `upto` is current"""
    for b in range(1, upto):
        ts.set_bond(b)
        ts.shiftE(W, b, "Fromleft")
    ts.set_bond(upto)


@pytest.mark.parametrize("b", [1, 2, 3, 4, 5, 6, 7])
def test_literal_equals_structured(b):
    feat, labels, W = make_problem(N=8, NT=60, m0=3)
    ts = O.TrainStates(feat, labels)
    ts.init(W)
    _walk(ts, W, b)
    B = O.form_bond(W[b], W[b + 1])
    P1 = O.project(B, ts, literal=True)
    P2 = O.project(B, ts, literal=False)
    assert rel(P1, P2) < 1e-12
    dP = np.random.default_rng(b).standard_normal(P1.shape)
    G1 = O.backproject(dP, B.shape, ts, literal=True)
    G2 = O.backproject(dP, B.shape, ts, literal=False)
    assert rel(G1, G2) < 1e-12
    # the env recursion equals the full contraction (util.h:19-40)
    for n in (0, 7, 33):
        assert rel(P2[n], O.toverlap(W, feat[n], ts.jc)) < 1e-11


@pytest.mark.parametrize("b", [2, 3, 4, 6])
def test_gradient_is_minus_half_dC(b):
    """finite-difference check: dC/dB = -2 * sum_n dP_n v_n"""
    feat, labels, W = make_problem(N=8, NT=40, m0=3)
    ts = O.TrainStates(feat, labels)
    ts.init(W)
    _walk(ts, W, b)
    B = O.form_bond(W[b], W[b + 1])
    G, _ = O._grad(B, ts, 0.0, False)
    rng = np.random.default_rng(1)
    for _ in range(3):
        dB = rng.standard_normal(B.shape)
        eps = 1e-6 * np.linalg.norm(B) / np.linalg.norm(dB)
        c1 = O.quadcost(B + eps * dB, ts)
        c0 = O.quadcost(B - eps * dB, ts)
        fd = (c1 - c0) / (2 * eps)
        an = -2.0 * float(np.sum(G * dB))
        assert abs(fd - an) < 1e-6 * max(abs(an), 1e-12) + 1e-9


def test_cg_decreases_quadratic_cost():
    feat, labels, W = make_problem(N=8, NT=200, m0=3)
    ts = O.TrainStates(feat, labels)
    ts.init(W)
    _walk(ts, W, 3)
    B0 = O.form_bond(W[3], W[4])
    c0 = O.quadcost(B0, ts)
    B, costs, rn = O.cgrad(B0, ts, 4)
    c4 = O.quadcost(B, ts)
    assert c4 < c0 and all(costs[i + 1] <= costs[i] + 1e-12 for i in range(len(costs) - 1))


@pytest.mark.parametrize("b,ha", [(2, 1), (3, 1), (4, 1), (4, 2), (3, 2), (6, 2)])
def test_svd_split_reconstructs(b, ha):
    feat, labels, W = make_problem(N=8, NT=50, m0=3)
    jc = 4
    rng = np.random.default_rng(b * 7 + ha)
    B = O.form_bond(W[b], W[b + 1])
    B = B + 0.1 * rng.standard_normal(B.shape)
    Wb, Wb1, m, te = O.svd_split(B, b, ha, jc, 100, 1, 0.0)
    assert rel(O.form_bond(Wb, Wb1), B) < 1e-12
    # the "c" side is an isometry
    if ha == 1:
        U = (np.transpose(Wb, (0, 1, 3, 2)) if Wb.ndim == 4 else Wb).reshape(-1, m) if Wb.ndim == 3 else \
            np.transpose(Wb, (0, 1, 3, 2)).reshape(-1, m)
        assert rel(U.T @ U, np.eye(m)) < 1e-12
    else:
        V = Wb1.reshape(m, -1)
        assert rel(V @ V.T, np.eye(m)) < 1e-12
    # truncation error equals the squared distance
    Wb, Wb1, m2, te = O.svd_split(B, b, ha, jc, 3, 1, 0.0)
    assert m2 == 3
    assert np.isclose(np.sum((O.form_bond(Wb, Wb1) - B) ** 2), te, rtol=1e-9)


def test_bond_position_invariance():
    """quadcost(newB) at bond b == quadcost(W(b+1)W(b+2)) at the next bond."""
    feat, labels, W = make_problem(N=10, NT=150, m0=3)
    ts = O.TrainStates(feat, labels)
    ts.init(W)
    rec = O.mldmrg(W, ts, 1, 8, 1, 1e-14, max_bonds=4)
    ts.set_bond(5)
    c = O.quadcost(O.form_bond(W[5], W[6]), ts) / ts.NT
    assert abs(c - rec[-1]["cost"]) < 1e-10 * max(1.0, abs(c))


def test_sharded_sweep_equals_single(tmp_path):
    feat, labels, W = make_problem(N=8, NT=120, m0=3)
    W2 = copy_mps(W)
    a = O.TrainStates(feat, labels, 1)
    a.init(W)
    b = O.TrainStates(feat, labels, 3)
    b.init(W2)
    ra = O.mldmrg(W, a, 1, 6, 1, 1e-12, max_bonds=3)
    rb = O.mldmrg(W2, b, 1, 6, 1, 1e-12, max_bonds=3)
    for x, y in zip(ra, rb):
        assert abs(x["cost"] - y["cost"]) < 1e-7 and x["m"] == y["m"]


@pytest.mark.parametrize("b,nthread", [(2, 1), (2, 3), (4, 2), (7, 2)])
def test_cpp_literal_port_matches_numpy_oracle(tmp_path, b, nthread):
    """oracle/fixedl_ref_cpu.cpp (the CPU baseline that bench.py times) and the
    numpy oracle are independent restatements; they must agree."""
    from oracle import cpu_ref
    feat, labels, W = make_problem(N=8, NT=900, m0=3)
    ts = O.TrainStates(feat, labels, nthread)
    ts.init(W)
    _walk(ts, W, b)
    B = O.form_bond(W[b], W[b + 1])
    Bo, costs, _ = O.cgrad(B, ts, 4)
    Co, _, ncor = O.quadcost(Bo, ts, detail=True)
    prob, res = str(tmp_path / "p.bin"), str(tmp_path / "r.bin")
    cpu_ref.write_problem(prob, *cpu_ref.problem_from_oracle(ts, B))
    out = cpu_ref.run(prob, res, nthread, 1, B.shape)
    assert rel(out["costs"], costs) < 1e-9
    # B itself is only comparable where the CG is well conditioned (the last step
    # a = |r|^2/pAp amplifies rounding along flat directions, DESIGN.md "Precision")
    if b != 4:
        assert rel(out["B"], Bo) < 1e-5
    # final cost: the numpy oracle run with 1/2/5 ParallelDo shards already spreads by
    # 4e-5 at the class-C bond (b=4); elsewhere summation order matters < 1e-9
    assert abs(out["C"] - Co) < (1e-3 if b == 4 else 1e-7) * Co and abs(out["ncor"] - ncor) <= 2


def _mps_dense(R):
    """Contract a raw MPS (list of [ml, p, mr], 1-indexed) to the dense tensor [p1, p2, ..., pN]."""
    T = R[1][0]                                   # [p, mr]
    for j in range(2, len(R)):
        T = np.tensordot(T, R[j], axes=([-1], [0]))
    return T[..., 0]


def test_mps_sum_is_exact_without_truncation():
    """Oracle of the initial-W construction (fixedL.cc:702-728): direct sum + orthogonalize with
    Cutoff 0 reproduces the dense sum of the terms; with Maxm the result is the best approximation
    bond by bond (error bounded by the discarded weight) and never exceeds Maxm."""
    rng = np.random.default_rng(4)
    N = 6

    def rand_mps(m):
        dims = [1] + [m] * (N - 1) + [1]
        return [None] + [rng.standard_normal((dims[j - 1], 2, dims[j])) for j in range(1, N + 1)]
    terms = [rand_mps(2), rand_mps(3), rand_mps(1)]
    dense = sum(_mps_dense(t) for t in terms)
    S = O.mps_sum([[None] + [a.copy() for a in t[1:]] for t in terms], 0.0, 10 ** 6)
    assert np.allclose(_mps_dense(S), dense, rtol=0, atol=1e-12 * np.abs(dense).max())
    assert abs(O.mps_overlap(S, S) - np.sum(dense * dense)) < 1e-10 * np.sum(dense * dense)
    T = O.mps_sum([[None] + [a.copy() for a in t[1:]] for t in terms], 0.0, 3)
    assert max(a.shape[2] for a in T[1:]) <= 3
    err = np.linalg.norm(_mps_dense(T) - dense) / np.linalg.norm(dense)
    assert 0 < err < 0.9                               # truncated, but still an approximation of the sum
    # product states: the sum of identical states is 2 x the state with bond dimension 1
    f = rng.uniform(0, 1, (N, 2))
    p = O.mps_product_state(f)
    P2 = O.mps_sum([p, O.mps_product_state(f)], 1e-10, 10)
    assert max(a.shape[2] for a in P2[1:]) == 1
    assert np.allclose(_mps_dense(P2), 2 * _mps_dense(p), atol=1e-12)


def test_ozaki_slicing_is_error_free_up_to_the_truncation():
    """oracle/ozaki.py (checker for the planned int8 tensor-core route): slices reconstruct the operand
    to 2^-7s, every slice fits a signed 7-bit integer, the product error falls by 2^-7 per slice and
    reaches float64 level at 8 slices."""
    from oracle import ozaki
    rng = np.random.default_rng(1)
    A = rng.standard_normal((64, 48)) * np.exp(3 * rng.standard_normal((64, 1)))
    B = rng.standard_normal((48, 40)) * np.exp(3 * rng.standard_normal((1, 40)))
    q, sc = ozaki.slices(A, 1, 6)
    assert all(np.abs(x).max() <= 127 for x in q)
    rec = sum(x * 2.0 ** (-7 * (i + 1)) for i, x in enumerate(q)) * sc
    # 6 slices = 42 bits below the power-of-two scale, which is at most 4x the row maximum
    assert np.max(np.abs(rec - A) / np.max(np.abs(A), axis=1, keepdims=True)) < 2.0 ** (-39)
    ref = A @ B
    bound = np.abs(A) @ np.abs(B)
    errs = [np.max(np.abs(ozaki.ozaki_matmul(A, B, s) - ref) / bound) for s in (4, 6, 8)]
    assert errs[0] < 1e-6 and errs[1] < 1e-10 and errs[2] < 1e-14
    assert errs[0] > 100 * errs[1] and errs[1] > 100 * errs[2]    # about 2^-14 per two slices


def test_oz_kernel_model_digits_and_accuracy():
    """oracle/ozaki.py::oz_kernel_model, the bit-level model of the tcgen05 kernel that was built: round-to-nearest
    digits through the 1.5 * 2^52 shift equal rint digits bit for bit (ties included), |digit| <= 64, the planes
    reconstruct the operand to 2^-57 of the row scale, and the modelled product agrees with the float64 contraction
    to a few 1e-16 of sum |a||b| at 8 planes, 1e-13 at 7, 1e-11 at 6 (the numbers tools/oz_test measures on the GPU)."""
    from oracle import ozaki
    rng = np.random.default_rng(2)
    A = rng.standard_normal((96, 120)) * np.exp(4 * rng.standard_normal((96, 1)))
    A[5] = 0.0                                                # an all-zero row: scale 1, digits 0
    A[7, :8] = A[7].max() * np.array([0.5, -0.5, 0.25, 1.0 / 256, -1.0 / 256, 3.0 / 256, 0.5 + 2.0 ** -8, 2.0 ** -20])   # ties
    qa, sa = ozaki.slices_rn(A, 1, 8, shift_trick=True)
    qb, sb = ozaki.slices_rn(A, 1, 8, shift_trick=False)
    assert all(np.array_equal(x, y) for x, y in zip(qa, qb)) and np.array_equal(sa, sb)
    assert max(int(np.abs(x).max()) for x in qa) <= 64
    rec = sum(x * 2.0 ** (-7 * (i + 1)) for i, x in enumerate(qa)) * sa
    assert np.max(np.abs(rec - A) / sa) <= 2.0 ** -57
    Bm = rng.standard_normal((240, 24)) * np.exp(2 * rng.standard_normal((1, 24)))
    f1 = np.stack([np.ones(96), rng.random(96) * 1e-2], axis=1)
    f1[3, 1] = 0.0
    ref = np.einsum("ra,apj,rp->rj", A, Bm.reshape(120, 2, 24), f1)
    bound = np.einsum("ra,apj,rp->rj", np.abs(A), np.abs(Bm.reshape(120, 2, 24)), np.abs(f1)) + 1e-300
    errs = [float(np.max(np.abs(ozaki.oz_kernel_model(A, Bm, f1, None, 2, ns) - ref) / bound)) for ns in (8, 7, 6)]
    assert errs[0] < 1e-15 and errs[1] < 1e-12 and errs[2] < 1e-10 and errs[1] > 10 * errs[0]
    # S = 4 (the projection: two feature pairs), one image per 3 rows
    f2 = np.stack([np.ones(32), rng.random(32)], axis=1)
    B4 = rng.standard_normal((480, 8))
    w4 = np.stack([f1[:32, 0] * f2[:, 0], f1[:32, 0] * f2[:, 1], f1[:32, 1] * f2[:, 0], f1[:32, 1] * f2[:, 1]], axis=1)
    ref4 = np.einsum("ra,apj,rp->rj", A, B4.reshape(120, 4, 8), np.repeat(w4, 3, axis=0))
    got4 = ozaki.oz_kernel_model(A, B4, f1[:32], f2, 4, 8, div=3)
    b4 = np.einsum("ra,apj,rp->rj", np.abs(A), np.abs(B4.reshape(120, 4, 8)), np.abs(np.repeat(w4, 3, axis=0))) + 1e-300
    assert float(np.max(np.abs(got4 - ref4) / b4)) < 1e-15


def test_golden_one_step_fixture_is_current():
    """tests/golden/config1_one_step.npz (golden OUTPUT vectors: cost, per-label cost, #correct, first CG pass,
    truncated bond tensor at 12 bonds of config 1's inputs) is what tests/golden/make_golden_steps.py produces
    from today's oracle -- a regression pin of the oracle; the GPU path is held to the same file."""
    import importlib.util
    here = os.path.join(os.path.dirname(__file__), "golden")
    spec = importlib.util.spec_from_file_location("make_golden_steps", os.path.join(here, "make_golden_steps.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    now = mod.compute()
    f = np.load(os.path.join(here, "config1_one_step.npz"))
    assert sorted(f.files) == sorted(now.keys())
    for k in f.files:
        a, b = np.asarray(f[k], np.float64), np.asarray(now[k], np.float64)
        assert a.shape == b.shape, k
        tol = 1e-7 if k.startswith("rn1_") else 1e-11
        assert np.abs(a - b).max() <= tol * max(np.abs(a).max(), 1e-300), k
    # the model function does not depend on the bond it is evaluated at
    C = [float(f[f"C{b}"][0]) for b in f["bonds"]]
    assert max(C) - min(C) < 1e-9 * C[0]


def test_sharded_oracle_equals_whole():
    """tests/helpers.ShardedOracle (the checker of the large-shape GPU tests) == the plain oracle."""
    from tests.helpers import ShardedOracle, make_problem, rel
    feat, labels, W = make_problem(N=10, NT=300, m0=4)
    ts = O.TrainStates(feat, labels)
    ts.init(W)
    so = ShardedOracle(feat, labels, chunk=128)
    so.init(W)
    for b in range(1, 7):
        ts.set_bond(b)
        so.set_bond(b)
        B = O.form_bond(W[b], W[b + 1])
        assert rel(so.project(B), O.project(B, ts)) < 1e-13
        C, CL, nc = O.quadcost(B, ts, detail=True)
        c, cl, n = so.quadcost(B)
        assert abs(c - C) < 1e-12 * C and n == nc and rel(cl, CL) < 1e-12
        Bo, costs, rn = O.cgrad(B, ts, 3, 1e-4)
        Bs, cs, rs, steps = so.cgrad(B, 3, 1e-4)
        B1, _, _ = O.cgrad(B, ts, 1, 1e-4)
        # sharp on one-step quantities only: the CG amplifies the shard-order rounding
        assert rel(cs[:1], costs[:1]) < 1e-12 and rel(steps[0], B1 - B) < 1e-10 and len(steps) == 3
        assert rel(Bs, Bo) < 1e-2
        ts.shiftE(W, b, "Fromleft")
        so.shiftE(W, b, "Fromleft")
        assert rel(so.slot(b), ts.slot[b]) < 1e-13
