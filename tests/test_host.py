"""CPU tests of the host side: the C-ABI library loads and exports every
declared symbol, fails loudly without a GPU, the input-file parser, shard
bounds, the drop-in binary's usage behaviour, and the N>1 plumbing over gloo."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_all_exported():
    from tnml_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "tnml_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(tnml_[a-z_]+)\s*\(", hdr)))
    declared = [d for d in declared if d not in ("tnml_handle_s",)]
    lib = capi.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(capi.SYMBOLS) == declared
    assert b"sm_100a" in lib.tnml_version()


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tnml_b200 import capi
    with pytest.raises(capi.TnmlError) as e:
        capi.Handle(0)
    assert e.value.code == -3 and "no CPU path" in str(e.value)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tnml_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f


def test_input_group_parser(tmp_path, ref_mnist_dir):
    from tnml_b200.data import InputGroup
    ig = InputGroup("/root/reference/sample_inputs/input_fixedL")
    assert ig.getInt("Ntrain", 60000) == 100 and ig.getInt("maxm", 5000) == 40
    assert ig.getReal("cutoff", 1e-10) == 1e-12 and ig.getReal("lambda", 0.0) == 1e-3
    assert ig.getInt("Npass", 4) == 2 and ig.getInt("nthread", 1) == 2 and ig.getInt("Nbatch", 10) == 4
    assert ig.getInt("imglen", 0) == 28               # read by us, ignored by the reference
    assert ig.getString("method", "conj") == "conj"    # default for an absent key
    assert ig.getYesNo("replace", False) is False


def test_readmnist_matches_oracle(ref_mnist_dir):
    from oracle import fixedl_oracle as O
    from tnml_b200 import data
    d1, l1 = data.readMNIST(ref_mnist_dir, "Train", 7)
    d2, l2, _ = O.read_mnist(ref_mnist_dir, "Train", 7)
    assert np.array_equal(d1, d2) and np.array_equal(l1, l2)
    assert np.allclose(data.reduce(d1, 14), O.reduce_image(d2, 14), rtol=0, atol=0)
    assert np.array_equal(data.phi(d1[:3]), O.features(d2[:3]))


def test_bounds_and_sweepnext():
    from tnml_b200 import fixedl
    assert fixedl.bounds(4, 10) == [(0, 2), (2, 4), (4, 6), (6, 10)]
    assert list(fixedl.sweepnext(4)) == [(1, 1), (2, 1), (3, 1), (3, 2), (2, 2), (1, 2)]


def _host_bin():
    p = os.path.join(ROOT, "tnml_b200", "host", "fixedL")
    if not os.path.exists(p):
        subprocess.run(["make", "-s", "host"], cwd=ROOT, check=True)
    return p


def test_host_selftest():
    """itensor_lite.h (contraction, addition, commonIndex, Sweeps, Args, sweepnext) and initial_w.h
    (small SVD, MPS direct sum + compression, overlap bilinearity, Maxm) -- C++ self-test, no GPU."""
    _host_bin()
    p = os.path.join(ROOT, "tnml_b200", "host", "hosttest")
    if not os.path.exists(p):
        subprocess.run(["make", "-s", "host"], cwd=ROOT, check=True)
    r = subprocess.run([p], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "hosttest: PASS" in r.stdout, r.stdout + r.stderr


def test_fixedL_usage_returns_zero():
    r = subprocess.run([_host_bin()], capture_output=True, text=True)
    assert r.returncode == 0 and "Usage:" in r.stdout and "inputfile" in r.stdout   # fixedL.cc:579-583


def test_fixedL_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tnml_b200 import data
    pix, labels = data.synthetic_digits(40, 14, seed=3)
    data.write_idx_files(str(tmp_path / "d"), (pix * 255).round().astype(np.uint8), labels, 14)
    (tmp_path / "in").write_text(f"input\n{{\ndatadir = {tmp_path}/d\nNtrain = 4\nimglen = 14\nNbatch = 4\nmaxm = 4\n}}\n")
    r = subprocess.run([_host_bin(), "in"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU path" in r.stderr
    # Nbatch must divide the image count (fixedL.cc:84-89)
    (tmp_path / "in2").write_text(f"input\n{{\ndatadir = {tmp_path}/d\nNtrain = 4\nimglen = 14\nNbatch = 7\n}}\n")
    r = subprocess.run([_host_bin(), "in2"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "not commensurate" in r.stderr


def test_initial_w_sum_matches_oracle(tmp_path):
    """the initial W (fixedL.cc:702-728: per label a compressed sum of `ninitial`
    random product states, tagged 0.1*setElt(L), the ten summed, centre normalised) built by the
    host program equals the oracle's restatement for the same drawn images, and equals the exact
    (uncompressed) sum of product states up to the truncation error.  No GPU needed: `init_only`."""
    from oracle import fixedl_oracle as O
    from tnml_b200 import data
    side, per, nini = 6, 12, 7
    N = side * side
    pix, labels = data.synthetic_digits(10 * per, side, seed=5)
    u8 = (pix * 255).round().astype(np.uint8)
    data.write_idx_files(str(tmp_path / "d"), u8, labels, side)
    (tmp_path / "in").write_text(f"input\n{{\ndatadir = {tmp_path}/d\nNtrain = {per}\nimglen = {side}\nNbatch = 4\n"
                                 f"ninitial = {nini}\nseed = 9\ninit_only = yes\n}}\n")
    r = subprocess.run([_host_bin(), "in"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Summing %d random label 3 states" % nini in r.stdout and "Summing all 10 label states together" in r.stdout
    picks = [[int(x) for x in re.search(r"label %d picks:([ 0-9]+)" % l, r.stdout).group(1).split()] for l in range(10)]
    assert all(len(p) == nini for p in picks)
    sel = np.array([labels[i] for i in range(len(labels))])
    for l in range(10):
        assert all(sel[i] == l for i in picks[l])             # randImg retries until the label matches
    feat = data.phi(u8.astype(np.float64) / 255.0)            # same double /255 as the host
    jc = N // 2
    Wh = data.read_mps_file(str(tmp_path / "W"))
    Wo = O.initial_w_sum(feat, picks, jc)
    assert Wh[jc].ndim == 4 and Wh[jc].shape[3] == 10 and abs(np.linalg.norm(Wh[jc]) - 1.0) < 1e-12

    def raw(W):   # label site -> [ml, d*NL, mr] so that the overlap runs over the label index too
        R = list(W)
        A = W[jc]
        R[jc] = np.transpose(A, (0, 1, 3, 2)).reshape(A.shape[0], -1, A.shape[2])
        return R
    hh, oo, ho = O.mps_overlap(raw(Wh), raw(Wh)), O.mps_overlap(raw(Wo), raw(Wo)), O.mps_overlap(raw(Wh), raw(Wo))
    assert abs(ho / np.sqrt(hh * oo) - 1.0) < 1e-9            # same state (gauge free)
    assert abs(hh / oo - 1.0) < 1e-9
    assert max(w.shape[2] for w in Wh[1:]) <= 10              # Maxm = 10
    # against the exact sum: W(x)_l  ~  0.1/norm * sum_{n in picks[l]} prod_j <phi(x_j)|phi(n_j)>
    x = feat[:40]
    Pw = np.array([O.toverlap(Wh, x[i], jc) for i in range(len(x))])
    ex = np.stack([np.prod(np.einsum("ijs,njs->inj", x, feat[picks[l]]), axis=2).sum(1) for l in range(10)], axis=1)
    scale = (Pw * ex).sum() / (ex * ex).sum()
    assert np.abs(Pw - scale * ex).max() < 1e-3 * np.abs(Pw).max()


def test_w0_to_w9_merge_matches_oracle(tmp_path):
    """fixedL.cc:682-701: separate label MPS W0..W9 are tagged with setElt(L(label)) on site c and
    summed with Cutoff 1E-10."""
    from oracle import fixedl_oracle as O
    from tnml_b200 import data
    side = 4
    N, jc = side * side, side * side // 2
    pix, labels = data.synthetic_digits(40, side, seed=2)
    data.write_idx_files(str(tmp_path / "d"), (pix * 255).round().astype(np.uint8), labels, side)
    data.write_sites_file(str(tmp_path / "sites"), N)
    rng = np.random.default_rng(0)
    terms = []
    for l in range(10):
        dims = [1] + [min(3, 2 ** min(j, N - j)) for j in range(1, N)] + [1]
        Wl = [None] + [rng.standard_normal((dims[j - 1], 2, dims[j])) / np.sqrt(dims[j - 1]) for j in range(1, N + 1)]
        data.write_mps_file(str(tmp_path / f"W{l}"), Wl)
        T = [None] + [w.copy() for w in Wl[1:]]
        A = T[jc]
        Z = np.zeros((A.shape[0], 2, 10, A.shape[2]))
        Z[:, :, l, :] = A
        T[jc] = Z.reshape(A.shape[0], 20, A.shape[2])
        terms.append(T)
    (tmp_path / "in").write_text(f"input\n{{\ndatadir = {tmp_path}/d\nNtrain = 4\nimglen = {side}\nNbatch = 4\n"
                                 f"init_only = yes\n}}\n")
    r = subprocess.run([_host_bin(), "in"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Found separate W0,W1,...,W9 MPS: summing" in r.stdout, r.stderr[-2000:]
    Wh = data.read_mps_file(str(tmp_path / "W"))
    Ro = O.mps_sum(terms, 1e-10, 10 ** 6)
    Rh = list(Wh)
    A = Wh[jc]
    Rh[jc] = np.transpose(A, (0, 1, 3, 2)).reshape(A.shape[0], -1, A.shape[2])
    hh, oo, ho = O.mps_overlap(Rh, Rh), O.mps_overlap(Ro, Ro), O.mps_overlap(Rh, Ro)
    assert abs(ho / np.sqrt(hh * oo) - 1.0) < 1e-10 and abs(hh / oo - 1.0) < 1e-10
    exact = sum(O.mps_overlap(a, b) for a in terms for b in terms)      # |sum_l psi_l (x) e_l|^2
    assert abs(hh / exact - 1.0) < 1e-8


WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["TNML_ROOT"])
import numpy as np, torch, torch.distributed as dist
from oracle import fixedl_oracle as O
from tnml_b200 import fixedl
from tests.helpers import make_problem
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
feat, labels, W = make_problem(N=8, NT=101, m0=3)
a, b = fixedl.bounds(world, 101)[rank]                 # ParallelDo shard of this rank
ts = O.TrainStates(feat[a:b], labels[a:b]); ts.init(W); ts.set_bond(1)
B = O.form_bond(W[1], W[2])
G, C = O._grad(B, ts, 0.0, False)                      # per-shard gradient and cost
buf = torch.from_numpy(np.concatenate([G.ravel(), [C]]))
dist.all_reduce(buf)                                   # the only exchange step of the path
if rank == 0:
    whole = O.TrainStates(feat, labels); whole.init(W); whole.set_bond(1)
    Gw, Cw = O._grad(B, whole, 0.0, False)
    err = np.abs(buf.numpy()[:-1] - Gw.ravel()).max() / np.abs(Gw).max()
    assert err < 1e-12 and abs(buf.numpy()[-1] - Cw) < 1e-10 * Cw, (err,)
    # the NCCL unique id travels as a 128-byte broadcast
uid = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    uid = torch.arange(128, dtype=torch.uint8)
dist.broadcast(uid, 0)
assert uid.tolist() == list(range(128))
dist.destroy_process_group()
print("ok", rank)
'''


def test_two_rank_gloo_gradient_allreduce(tmp_path):
    """N>1 path on CPU: images sharded with ParallelDo bounds, per-shard gradient +
    cost all-reduced (gloo stands in for NCCL), result equals the single-rank
    gradient; the unique-id broadcast plumbing of bench.py works."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, TNML_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29731", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
