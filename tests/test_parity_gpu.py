"""Parity tests proper: the CUDA path, called through the C-ABI, against the
float64 oracle on the same seeded inputs.  Tolerances (all float64):
  forward quantities (P, cost, envs, gradient)      rel 1e-11
  B after the 4-pass CG                             rel 1e-6   (the CG itself
      amplifies 1e-16 reordering noise to ~1e-7 between two float64 oracles,
      see DESIGN.md "Precision")
  cost after a whole bond update                    rel 1e-8
"""
import numpy as np
import pytest

from oracle import fixedl_oracle as O
from tests.helpers import copy_mps, make_problem, rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from tnml_b200 import capi as c
    c.load_library()
    return c


def _gpu_state(capi, feat, labels, W):
    h = capi.Handle(0)
    h.set_images(feat, labels.astype(np.int32))
    h.set_mps(W)
    h.init_envs()
    return h


def _walk_both(h, ts, W, capi, upto):
    for b in range(1, upto):
        ts.set_bond(b)
        ts.shiftE(W, b, "Fromleft")
        h.set_bond(b)
        h.shift_env(b, capi.FROMLEFT)
    ts.set_bond(upto)
    h.set_bond(upto)


def test_version_and_errors(capi):
    assert b"sm_100a" in capi.load_library().tnml_version()
    h = capi.Handle(0)
    with pytest.raises(capi.TnmlError):
        h.init_envs()                      # no images
    feat, labels, W = make_problem(N=8, NT=20, m0=2)
    h.set_images(feat, labels.astype(np.int32))
    with pytest.raises(capi.TnmlError):
        h.set_site(3, W[4])                # label index on the wrong site (fixedL.cc:734)
    with pytest.raises(capi.TnmlError):
        h.init_envs()                      # sites missing
    h.close()


def test_init_envs_match_oracle(capi):
    feat, labels, W = make_problem(N=10, NT=257, m0=4)
    ts = O.TrainStates(feat, labels)
    ts.init(W)
    h = _gpu_state(capi, feat, labels, W)
    for j in range(3, 11):
        assert rel(h.get_env(j), ts.slot[j]) < 1e-12, j
    h.close()


@pytest.mark.parametrize("b", [1, 2, 3, 4, 5, 6, 7, 8, 9])
def test_forward_and_gradient_all_bond_classes(capi, b):
    """classes L (b<=3), C (b=4,5), R (b>=6) incl. the edge bonds (ml=1 / mr=1)."""
    feat, labels, W = make_problem(N=10, NT=333, m0=4)
    ts = O.TrainStates(feat, labels)
    ts.init(W)
    h = _gpu_state(capi, feat, labels, W)
    _walk_both(h, ts, W, capi, b)
    if b > 1:
        assert rel(h.get_env(b - 1), ts.slot[b - 1]) < 1e-12
    B = O.form_bond(W[b], W[b + 1])
    h.bond_form()
    assert rel(h.bond_store(), B) < 1e-13
    C, CL, ncor = O.quadcost(B, ts, detail=True)
    c, cl, nc = h.quadcost(False)
    assert abs(c - C) < 1e-11 * C and rel(cl, CL) < 1e-11 and nc == ncor
    lab, P = h.predict(want_P=True)
    Pref = O.project(B, ts)
    assert rel(P, Pref) < 1e-11
    assert np.array_equal(lab, O.argmax_first(np.abs(Pref)))
    # one CG pass == gradient + step: checks backward too
    Bo, costs_o, rn_o = O.cgrad(B, ts, 2)
    h.bond_load(B)
    costs, rn = h.cgrad(2)
    assert rel(h.bond_store(), Bo) < 1e-9
    assert rel(costs, costs_o) < 1e-10 and rel(rn, rn_o) < 1e-7
    h.close()


@pytest.mark.parametrize("lam", [0.0, 1e-3])
def test_cgrad_matches_oracle(capi, lam):
    feat, labels, W = make_problem(N=10, NT=500, m0=4)
    ts = O.TrainStates(feat, labels)
    ts.init(W)
    h = _gpu_state(capi, feat, labels, W)
    _walk_both(h, ts, W, capi, 3)
    B = O.form_bond(W[3], W[4])
    Bo, costs_o, rn_o = O.cgrad(B, ts, 4, lam)
    h.bond_load(B)
    costs, rn = h.cgrad(4, lam)
    assert len(costs) == len(costs_o)
    assert rel(costs, costs_o) < 1e-9
    assert rel(h.bond_store(), Bo) < 1e-6
    assert abs(h.quadcost(False, lam)[0] - O.quadcost(Bo, ts, lam)) < 1e-9 * O.quadcost(Bo, ts, lam)
    h.close()


@pytest.mark.parametrize("b,ha", [(2, 1), (4, 1), (5, 1), (5, 2), (4, 2), (8, 2), (9, 1), (1, 2)])
def test_svd_split_matches_oracle(capi, b, ha):
    """gauge-free comparison: newB = W(c)W(c+dc), m, truncerr; isometry of W(c)."""
    feat, labels, W = make_problem(N=10, NT=64, m0=4)
    ts = O.TrainStates(feat, labels)
    ts.init(W)
    h = _gpu_state(capi, feat, labels, W)
    _walk_both(h, ts, W, capi, b)
    rng = np.random.default_rng(100 * b + ha)
    B = O.form_bond(W[b], W[b + 1])
    B = B + 0.05 * np.linalg.norm(B) / np.sqrt(B.size) * rng.standard_normal(B.shape)
    for (maxm, minm, cutoff) in [(100, 1, 0.0), (5, 1, 0.0), (6, 3, 1e-3)]:
        Wb, Wb1, m, te = O.svd_split(B, b, ha, 5, maxm, minm, cutoff)
        h.bond_load(B)
        gm, gte = h.svd_split(capi.FROMLEFT if ha == 1 else capi.FROMRIGHT, cutoff, maxm, minm)
        assert gm == m
        assert abs(gte - te) <= 1e-9 * max(te, 1e-300) + 1e-24
        gWb, gWb1 = h.get_site(b), h.get_site(b + 1)
        assert gWb.shape == Wb.shape and gWb1.shape == Wb1.shape
        assert rel(O.form_bond(gWb, gWb1), O.form_bond(Wb, Wb1)) < 1e-11
        iso = gWb if ha == 1 else gWb1
        if ha == 1:
            U = (np.transpose(iso, (0, 1, 3, 2)) if iso.ndim == 4 else iso).reshape(-1, m)
            assert rel(U.T @ U, np.eye(m)) < 1e-12
        else:
            V = iso.reshape(m, -1)
            assert rel(V @ V.T, np.eye(m)) < 1e-12
    h.close()


def test_bond_update_sequence_matches_oracle(capi):
    """Whole loop body of mldmrg, bond after bond, on a chain that covers all
    three bond classes and both sweep directions.

    Yardstick: the reference algorithm is chaotic -- a single bond update can
    amplify 1e-16 summation-order noise by 1e9 (CG step a = |r|^2/pAp along
    nearly flat directions, DESIGN.md "Precision").  So the oracle is run twice
    (1 and 3 ParallelDo shards = two float64 summation orders) and the CUDA
    path must stay within 10x of the oracle-vs-oracle spread (floor 1e-9)."""
    feat, labels, W = make_problem(N=10, NT=1000, m0=3)
    refs = []
    for ns in (1, 3):
        ts = O.TrainStates(feat, labels, ns)
        ts.init(copy_mps(W))
        Wo = copy_mps(W)
        refs.append((O.mldmrg(Wo, ts, 1, 8, 4, 1e-10), Wo))
    ref, Wo = refs[0]
    h = _gpu_state(capi, feat, labels, W)
    p = capi.BondParams(4, 0.0, 1e-10, 1e-10, 8, 4, 0)
    worst, noise = 0.0, 0.0
    for k, (b, ha) in enumerate(O.sweep_schedule(10)):
        r = h.bond_update(b, ha, p)
        o = ref[k]
        noise = max(noise, abs(refs[1][0][k]["cost"] - o["cost"]) / o["cost"])
        assert r.newm == o["m"], (k, b, ha)
        e = abs(r.cost / 1000 - o["cost"]) / o["cost"]
        worst = max(worst, e)
        assert e < max(1e-9, 10 * noise), (k, b, ha, e, noise)
        assert abs(int(r.ncorrect) - o["ncor"]) <= 10      # argmax flips of near-tied outputs (1% of NT)
    print("worst rel cost deviation over the sweep:", worst, "oracle self-noise:", noise)
    # final MPS: compare the model function, not the gauge
    Wg = h.get_mps()
    for n in (0, 10, 500):
        assert rel(O.toverlap(Wg, feat[n], 5), O.toverlap(Wo, feat[n], 5)) < 1e-2
    h.close()


def test_golden_mnist_first_bonds(capi):
    """Real MNIST (committed 1000-image 14x14 subset): first 12 bond updates of
    BASELINE config 1 (maxm=20) against the oracle."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mnist_100_per_label_14x14.npz"))
    feat = O.features(g["sum4"].astype(np.float64) / (4 * 255.0))
    labels = g["labels"]
    from tnml_b200 import data as D
    W = D.random_mps(196, 2, 10, seed=1)
    refs = []
    for ns in (1, 4):     # two float64 summation orders; their spread is the yardstick
        ts = O.TrainStates(feat, labels, ns)
        ts.init(copy_mps(W))
        refs.append(O.mldmrg(copy_mps(W), ts, 1, 20, 10, 1e-10, max_bonds=12))
    ref = refs[0]
    h = _gpu_state(capi, feat, labels, W)
    p = capi.BondParams(4, 0.0, 1e-10, 1e-10, 20, 10, 0)
    noise = 0.0
    for k in range(12):
        r = h.bond_update(k + 1, 1, p)
        noise = max(noise, abs(refs[1][k]["cost"] - ref[k]["cost"]) / ref[k]["cost"])
        assert r.newm == ref[k]["m"]
        e = abs(r.cost / 1000 - ref[k]["cost"]) / ref[k]["cost"]
        assert e < max(1e-9, 10 * noise), (k, e, noise)
    h.close()


def test_golden_one_step_fixture(capi):
    """Committed golden OUTPUT vectors (tests/golden/config1_one_step.npz, written by make_golden_steps.py from the
    oracle on the golden MNIST subset): cost, per-label cost, #correct, cost and |r| after the first CG pass, link
    dimension and rebuilt bond tensor of the truncated SVD at 12 bonds (edge bonds, class L, both class-C bonds,
    class R) of the rightward walk with W unchanged.  No oracle code runs here: the CUDA path meets the file."""
    import os
    here = os.path.join(os.path.dirname(__file__), "golden")
    g = np.load(os.path.join(here, "mnist_100_per_label_14x14.npz"))
    f = np.load(os.path.join(here, "config1_one_step.npz"))
    feat = O.features(g["sum4"].astype(np.float64) / (4 * 255.0))
    labels = g["labels"]
    from tnml_b200 import data as D
    W = D.random_mps(196, 2, 10, seed=1)
    chk = sum(float(np.sum(w * w)) for w in W if w is not None)
    assert abs(chk - float(f["w_checksum"][0])) < 1e-12 * chk           # same seeded MPS as the generator used
    bonds = set(int(b) for b in f["bonds"])
    h = _gpu_state(capi, feat, labels, W)
    for b in range(1, 196):
        h.set_bond(b)
        if b in bonds:
            h.bond_form()
            c, cl, nc = h.quadcost(False)
            C = float(f[f"C{b}"][0])
            assert abs(c - C) < 1e-11 * C, b
            assert rel(cl, f[f"CL{b}"]) < 1e-10, b
            assert abs(int(nc) - int(f[f"ncor{b}"][0])) <= 2, b
            costs, rn = h.cgrad(2)
            c1, r1 = float(f[f"cost1_{b}"][0]), float(f[f"rn1_{b}"][0])
            assert abs(costs[0] - c1) < 1e-9 * c1, b
            assert abs(rn[0] - r1) < 1e-5 * r1, (b, rn[0], r1)
            h.bond_form()                                               # B = W(b) W(b+1) again
            gm, _ = h.svd_split(capi.FROMLEFT, 1e-10, 20, 10)
            assert gm == int(f[f"m{b}"][0]), b
            assert rel(O.form_bond(h.get_site(b), h.get_site(b + 1)), f[f"newB{b}"]) < 1e-10, b
            h.set_site(b, W[b])                                         # back to the walk's gauge
            h.set_site(b + 1, W[b + 1])
            h.set_bond(b)
        h.shift_env(b, capi.FROMLEFT)
    h.close()


def test_shards_sum_to_whole(capi):
    """Size-independent property used for multi-GPU: per-shard costs and
    gradients are plain sums over images (fixedL.cc:385,402,421)."""
    feat, labels, W = make_problem(N=10, NT=600, m0=4)
    hs = []
    for (a, b) in O.shard_bounds(3, 600):
        hs.append(_gpu_state(capi, feat[a:b], labels[a:b], W))
    whole = _gpu_state(capi, feat, labels, W)
    tot = 0.0
    ncs = 0
    for h in hs:
        h.set_bond(1)
        h.bond_form()
        c, _, nc = h.quadcost(False)
        tot += c
        ncs += nc
    whole.set_bond(1)
    whole.bond_form()
    c, _, nc = whole.quadcost(False)
    assert abs(c - tot) < 1e-11 * c and nc == ncs
    for h in hs + [whole]:
        h.close()


def test_fixedL_binary_matches_capi(capi, tmp_path):
    """The drop-in `fixedL <inputfile>` program (host C++ over the C-ABI) reproduces,
    log line for log line, the costs of the Python-driven path from the same W."""
    import os
    import re
    import subprocess
    from tnml_b200 import data as D
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    binp = os.path.join(root, "tnml_b200", "host", "fixedL")
    if not os.path.exists(binp):
        pytest.skip("host binary not built")
    side = 6
    pix, labels = D.synthetic_digits(200, side, seed=11)
    u8 = (pix * 255).round().astype(np.uint8)
    # per-label cap 20 keeps all 200 images (20 per label), file order
    D.write_idx_files(str(tmp_path / "d"), u8, labels, side)
    W = D.random_mps(side * side, 2, 4, seed=4)
    D.write_sites_file(str(tmp_path / "sites"), side * side)
    D.write_mps_file(str(tmp_path / "W"), W)
    (tmp_path / "in").write_text(
        f"input\n{{\ndatadir = {tmp_path}/d\nNtrain = 20\nimglen = {side}\nNbatch = 4\nmaxm = 6\nminm = 3\n"
        f"cutoff = 1E-10\nNsweep = 1\nNpass = 3\nnthread = 2\ntrace = trace.jsonl\ntrace_phases = yes\n}}\n")
    r = subprocess.run([binp, "in"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    costs = [float(x) for x in re.findall(r"--> After SVD, Cost = ([0-9.eE+-]+)", r.stdout)]
    ms = [int(x) for x in re.findall(r"Original m=\d+, New m=(\d+)", r.stdout)]
    assert len(costs) == 2 * (side * side - 1)
    feat = D.phi(u8.astype(np.float64) / 255.0)
    h = _gpu_state(capi, feat, labels.astype(np.int64), W)
    p = capi.BondParams(3, 0.0, 1e-10, 1e-10, 6, 3, 0)
    for k, (b, ha) in enumerate(O.sweep_schedule(side * side)):
        res = h.bond_update(b, ha, p)
        assert res.newm == ms[k]
        assert abs(res.cost / 200 - costs[k]) < 2e-10 + 1e-9 * costs[k], k    # printed with 10 decimals
    assert "Before starting DMRG Cost" in r.stdout and "Writing W to disk" in r.stdout
    # per-bond JSONL trace: one parseable line per bond update, same numbers as the log
    import json
    tr = [json.loads(l) for l in (tmp_path / "trace.jsonl").read_text().splitlines()]
    assert len(tr) == len(costs)
    for k, t in enumerate(tr):
        assert t["newm"] == ms[k] and abs(t["cost"] - costs[k]) < 1e-9 and t["launches"] > 0 and t["wall_ms"] > 0
        assert set(t["phase_ms"]) == {"proj", "grad", "fat", "svd", "shift", "other"} and t["phase_ms"]["svd"] > 0
    # the trained W written by fixedL is then evaluated by the drop-in `fulltest` program
    D.write_idx_files(str(tmp_path / "d"), u8, labels, side, kind="t10k")
    ft = os.path.join(root, "tnml_b200", "host", "fulltest")
    (tmp_path / "in_test").write_text(f"input\n{{\ndatadir = {tmp_path}/d\nimglen = {side}\n}}\n")
    r2 = subprocess.run([ft, "in_test"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r2.returncode == 0, r2.stderr[-2000:]
    mt = re.search(r"(\d+)/200 correct", r2.stdout)
    assert mt and abs(int(mt.group(1)) - int(res.ncorrect)) <= 2     # same images, same final W
    assert "Total # test images = 200" in r2.stdout
    h.close()


@pytest.mark.parametrize("m0,NT", [(6, 700), (7, 1100), (13, 2500)])
def test_krgemm_variants_agree(capi, m0, NT):
    """The persistent projection kernel (output-side Khatri-Rao weights, cp.async A tiles, resident
    B panel; used from 512 rows on) against the register-staged kernel and the oracle: environment
    advance (S=2), projection / cost (S=4) on every bond class, even and odd link dimensions
    (16-byte and 8-byte cp.async paths)."""
    feat, labels, W = make_problem(N=12, NT=NT, m0=m0)
    ts = O.TrainStates(feat, labels)
    ts.init(copy_mps(W))
    res = {}
    for variant in (1, 2):
        h = capi.Handle(0)
        h.set_option("krgemm_variant", variant)
        h.set_images(feat, labels.astype(np.int32))
        h.set_mps(W)
        h.init_envs()
        out = []
        for b in range(1, 12):
            h.set_bond(b)
            h.bond_form()
            C, CL, nc = h.quadcost(False, 0.0)
            out.append((C, nc))
            h.shift_env(b, capi.FROMLEFT)
        res[variant] = (out, [h.get_env(j) for j in (3, 5, 6, 9)])
        h.set_option("krgemm_variant", 2)
        h.close()
    for (c1, n1), (c2, n2) in zip(res[1][0], res[2][0]):
        assert abs(c1 - c2) <= 1e-12 * abs(c1) and n1 == n2
    for e1, e2 in zip(res[1][1], res[2][1]):
        assert rel(e2, e1) < 1e-13
    for b in range(1, 12):          # and against the oracle
        ts.set_bond(b)
        Co, _, nco = O.quadcost(O.form_bond(W[b], W[b + 1]), ts, detail=True)
        assert abs(res[2][0][b - 1][0] - Co) <= 1e-11 * Co and res[2][0][b - 1][1] == nco
        ts.shiftE(W, b, "Fromleft")


def test_fat_and_krgram_variants_agree(capi):
    """The bulk-async-copy (TMA) label-environment kernel and the cp.async gradient kernel against
    their register-staged predecessors and the oracle, on ONE-STEP quantities (a whole bond update is
    already chaotic on the class-C bond of this chain: three float64 summation orders of the oracle
    itself differ by 4e-3 in the cost): forward cost, first CG step a*p (gradient
    contraction + pAp pass), cost after the first step -- on a class-L, a class-C and a class-R bond."""
    feat, labels, W = make_problem(N=12, NT=1500, m0=6)
    lam = 1e-5
    ts = O.TrainStates(feat, labels)
    ts.init(copy_mps(W))
    hs = {}
    for variant in (1, 2):
        hs[variant] = _gpu_state(capi, feat, labels, W)
    for b in range(1, 9):
        ts.set_bond(b)
        B = O.form_bond(W[b], W[b + 1])
        if b in (3, 5, 8):
            B1o, _, _ = O.cgrad(B, ts, 1, lam)
            _, costs_o, _ = O.cgrad(B, ts, 2, lam)
            Co = O.quadcost(B, ts, lam)
            res = {}
            for variant in (1, 2):
                h = hs[variant]
                h.set_option("fat_variant", variant)
                h.set_option("krgram_variant", variant)
                h.set_bond(b)
                h.bond_load(B)
                c = h.quadcost(False, lam)[0]
                h.cgrad(1, lam)
                step = h.bond_store() - B
                h.bond_load(B)
                costs, _ = h.cgrad(2, lam)
                res[variant] = (c, step, costs[0])
                assert abs(c - Co) < 1e-11 * Co, (b, variant)
                assert rel(step, B1o - B) < 1e-9, (b, variant)
                assert abs(costs[0] - costs_o[0]) < 1e-10 * costs_o[0], (b, variant)
            assert abs(res[1][0] - res[2][0]) < 1e-12 * Co
            assert rel(res[1][1], res[2][1]) < 1e-9
        ts.shiftE(W, b, "Fromleft")
        for variant in (1, 2):
            hs[variant].set_bond(b)
            hs[variant].shift_env(b, capi.FROMLEFT)
    for variant in (1, 2):
        hs[variant].set_option("fat_variant", 1)
        hs[variant].set_option("krgram_variant", 2)
        hs[variant].close()


def test_svd_variants_agree(capi):
    """The SVD variants (cluster-resident Jacobi with cross-only inner tournaments = default; full
    tournaments; multi-launch Jacobi; one QR instead of sort + two QRs; no preconditioner) give the
    same truncated factorisation: same m, same truncation error, same U*S*V to 1e-11, isometry."""
    feat, labels, W = make_problem(N=16, NT=64, m0=20)
    h = _gpu_state(capi, feat, labels, W)
    b = 6                                   # class-L bond, 40 x 40 bond matrix (label site is 8)
    for bb in range(1, b):
        h.set_bond(bb)
        h.shift_env(bb, capi.FROMLEFT)
    h.set_bond(b)
    rng = np.random.default_rng(3)
    B0 = O.form_bond(W[b], W[b + 1])
    B = B0 + 1e-4 * np.linalg.norm(B0) / np.sqrt(B0.size) * rng.standard_normal(B0.shape)
    variants = [dict(), dict(svd_cross=0), dict(svd_cluster=0), dict(svd_precond=1), dict(svd_precond=0, svd_cluster=0)]
    outs = []
    for v in variants:
        for k2, val in v.items():
            h.set_option(k2, val)
        h.bond_load(B)
        m, te = h.svd_split(capi.FROMLEFT, 1e-10, 20, 10)
        Wb, Wb1 = h.get_site(b), h.get_site(b + 1)
        outs.append((m, te, O.form_bond(Wb, Wb1), Wb))
        for k2 in v:
            h.set_option(k2, -1)
    m0, te0, nb0, _ = outs[0]
    Wo, Wo1, mo, teo = O.svd_split(B, b, 1, 8, 20, 10, 1e-10)
    assert m0 == mo and abs(te0 - teo) <= 1e-8 * teo + 1e-22
    assert rel(nb0, O.form_bond(Wo, Wo1)) < 1e-10
    for m, te, nb, Wb in outs:
        assert m == m0 and abs(te - te0) <= 1e-8 * te0 + 1e-22
        assert rel(nb, nb0) < 1e-11
        U = Wb.reshape(-1, m)
        assert np.abs(U.T @ U - np.eye(m)).max() < 1e-10
    h.close()


def test_env_tier_bit_identical(capi):
    """environment tiering.  With an HBM budget that holds only a handful of
    environment slots (the rest live in pinned host memory, fetched ahead on a copy stream) two
    sweeps produce bit-identical costs, link dims and MPS as the all-resident run; the tier really
    moved data; an evicted slot can still be read back and equals the oracle's."""
    feat, labels, W = make_problem(N=16, NT=256, m0=4)
    p = capi.BondParams(3, 0.0, 1e-10, 1e-10, 8, 4, 0)
    runs = []
    for budget_gb in (0.0, 0.0012):      # 0.0012 GiB = 1.29 MB: ~7 label-carrying slots of 256x10x8 doubles
        h = capi.Handle(0)
        h.set_images(feat, labels.astype(np.int32))
        h.set_mps(W)
        if budget_gb:
            h.set_option("env_budget_gb", budget_gb)
        h.init_envs()
        h.stats(reset=True)
        out = []
        for sw in range(2):
            for b, ha in O.sweep_schedule(16):
                r = h.bond_update(b, ha, p)
                out.append((r.cost, r.newm, r.ncorrect, r.truncerr))
        st = h.stats(reset=True)
        mps = [h.get_site(j) for j in range(1, 17)]
        envs = {}
        for j in (1, 5, 9, 14):           # left envs far behind the last bond (bond 1, moving left): some evicted
            try:
                envs[j] = h.get_env(j)
            except capi.TnmlError:
                envs[j] = None
        runs.append((out, mps, st, envs))
        h.close()
    (o0, m0, s0, e0), (o1, m1, s1, e1) = runs
    assert o0 == o1                                           # bit-identical, not just close
    assert all(np.array_equal(a, b) for a, b in zip(m0, m1))
    assert s0.tier_evictions == 0 and s0.tier_fetches == 0
    assert s1.tier_evictions > 10 and s1.tier_fetches > 10 and s1.tier_bytes > 0
    for j in e0:
        if e0[j] is None:
            assert e1[j] is None
        else:
            assert np.array_equal(e0[j], e1[j])


def test_fixedL_cold_start_matches_oracle(capi, tmp_path):
    """Cold start of the drop-in program (no `W` file): the initial W is the reference's sum of
    product states (fixedL.cc:702-728); the cost before DMRG and the first bond
    updates equal the oracle's from the W the program wrote."""
    import os
    import re
    import subprocess
    from tnml_b200 import data as D
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    binp = os.path.join(root, "tnml_b200", "host", "fixedL")
    if not os.path.exists(binp):
        pytest.skip("host binary not built")
    side = 6
    N = side * side
    pix, labels = D.synthetic_digits(200, side, seed=21)
    u8 = (pix * 255).round().astype(np.uint8)
    D.write_idx_files(str(tmp_path / "d"), u8, labels, side)
    base = (f"datadir = {tmp_path}/d\nNtrain = 20\nimglen = {side}\nNbatch = 4\nmaxm = 8\nminm = 4\ncutoff = 1E-10\n"
            f"Nsweep = 1\nNpass = 3\nninitial = 6\nseed = 3\n")
    (tmp_path / "in0").write_text("input\n{\n" + base + "init_only = yes\n}\n")
    r0 = subprocess.run([binp, "in0"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r0.returncode == 0, r0.stderr[-2000:]
    W0 = D.read_mps_file(str(tmp_path / "W"))          # the initial W, before training overwrites it
    os.remove(tmp_path / "W")
    (tmp_path / "in").write_text("input\n{\n" + base + "}\n")
    r = subprocess.run([binp, "in"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Summing all 10 label states together" in r.stdout
    c0 = float(re.search(r"Before starting DMRG Cost = ([0-9.eE+-]+)", r.stdout).group(1))
    costs = [float(x) for x in re.findall(r"--> After SVD, Cost = ([0-9.eE+-]+)", r.stdout)]
    assert len(costs) == 2 * (N - 1)
    feat = D.phi(u8.astype(np.float64) / 255.0)
    lab = labels.astype(np.int64)
    ts = O.TrainStates(feat, lab)
    Wc = [None if w is None else w.copy() for w in W0]
    ts.init(Wc)
    C, _, _ = O.quadcost(O.form_bond(Wc[1], Wc[2]), ts, detail=True)
    assert abs(C / 200 - c0) < 2e-10 + 1e-9 * c0
    ref = O.mldmrg(Wc, ts, 1, 8, 4, 1e-10, Npass=3, max_bonds=6)
    for k in range(6):
        assert abs(ref[k]["cost"] - costs[k]) < 2e-10 + 1e-6 * costs[k], (k, ref[k]["cost"], costs[k])
    assert costs[-1] < c0                                # training reduced the cost


@pytest.mark.parametrize("b,ha", [(6, 1), (7, 1), (8, 1), (8, 2), (7, 2), (10, 2)])
def test_svd_qr_preconditioned_path(capi, b, ha):
    """Larger bond matrices (>= 32 columns) go through Householder QR + Jacobi on R^T
    + apply-Q; rank-deficient + noise spectra like a freshly optimised bond tensor."""
    feat, labels, W = make_problem(N=16, NT=64, m0=20)
    ts = O.TrainStates(feat, labels)
    ts.init(W)
    h = _gpu_state(capi, feat, labels, W)
    _walk_both(h, ts, W, capi, b)
    rng = np.random.default_rng(b * 10 + ha)
    B0 = O.form_bond(W[b], W[b + 1])
    for noise in (1e-1, 1e-6):
        B = B0 + noise * np.linalg.norm(B0) / np.sqrt(B0.size) * rng.standard_normal(B0.shape)
        for (maxm, minm, cutoff) in [(1000, 1, 0.0), (20, 10, 1e-10)]:
            Wb, Wb1, m, te = O.svd_split(B, b, ha, 8, maxm, minm, cutoff)
            h.bond_load(B)
            gm, gte = h.svd_split(capi.FROMLEFT if ha == 1 else capi.FROMRIGHT, cutoff, maxm, minm)
            assert gm == m
            assert abs(gte - te) <= 1e-8 * max(te, 1e-300) + 1e-22 * np.linalg.norm(B) ** 2
            gWb, gWb1 = h.get_site(b), h.get_site(b + 1)
            assert rel(O.form_bond(gWb, gWb1), O.form_bond(Wb, Wb1)) < 1e-10
            iso = gWb if ha == 1 else gWb1
            if ha == 1:
                U = (np.transpose(iso, (0, 1, 3, 2)) if iso.ndim == 4 else iso).reshape(-1, m)
                assert rel(U.T @ U, np.eye(m)) < 1e-12
            else:
                V = iso.reshape(m, -1)
                assert rel(V @ V.T, np.eye(m)) < 1e-12
            h.set_site(b, W[b])
            h.set_site(b + 1, W[b + 1])
    h.close()


@pytest.mark.parametrize("b,ha", [(7, 1), (8, 2), (7, 2)])
def test_svd_tall_class_c(capi, b, ha):
    """Class-C bond matrices at larger link dimension (1280-1400 x 128-140): the first QR and the
    apply-Q of the tall side run from shared-memory-resident columns (> 1280 rows)."""
    feat, labels, W = make_problem(N=16, NT=32, m0=70)
    ts = O.TrainStates(feat, labels)
    ts.init(W)
    h = _gpu_state(capi, feat, labels, W)
    _walk_both(h, ts, W, capi, b)
    rng = np.random.default_rng(b * 10 + ha)
    B0 = O.form_bond(W[b], W[b + 1])
    B = B0 + 1e-3 * np.linalg.norm(B0) / np.sqrt(B0.size) * rng.standard_normal(B0.shape)
    assert max(B.shape[0], B.shape[3]) * 2 * 10 > 1280
    for (maxm, minm, cutoff) in [(70, 35, 1e-10), (1000, 1, 0.0)]:
        Wb, Wb1, m, te = O.svd_split(B, b, ha, 8, maxm, minm, cutoff)
        h.bond_load(B)
        gm, gte = h.svd_split(capi.FROMLEFT if ha == 1 else capi.FROMRIGHT, cutoff, maxm, minm)
        assert gm == m
        assert abs(gte - te) <= 1e-8 * max(te, 1e-300) + 1e-22 * np.linalg.norm(B) ** 2
        gWb, gWb1 = h.get_site(b), h.get_site(b + 1)
        assert rel(O.form_bond(gWb, gWb1), O.form_bond(Wb, Wb1)) < 1e-10
        iso = gWb if ha == 1 else gWb1
        if ha == 1:
            U = (np.transpose(iso, (0, 1, 3, 2)) if iso.ndim == 4 else iso).reshape(-1, m)
            assert rel(U.T @ U, np.eye(m)) < 1e-11
        else:
            V = iso.reshape(m, -1)
            assert rel(V @ V.T, np.eye(m)) < 1e-11
        h.set_site(b, W[b])
        h.set_site(b + 1, W[b + 1])
    h.close()


@pytest.mark.parametrize("b", [3, 4, 7])
def test_cg_reuse_forward_option(capi, b):
    """cg_reuse_forward=1 (linear update of the forward outputs) is the same mathematics as the
    literal recompute; results agree to the CG's own noise amplification (classes L, C, R)."""
    feat, labels, W = make_problem(N=10, NT=700, m0=4)
    ts = O.TrainStates(feat, labels)
    ts.init(W)
    h = _gpu_state(capi, feat, labels, W)
    _walk_both(h, ts, W, capi, b)
    B = O.form_bond(W[b], W[b + 1])
    Bo, costs_o, _ = O.cgrad(B, ts, 4)
    h.set_option("cg_reuse_forward", 0)
    h.bond_load(B)
    c0, _ = h.cgrad(4)
    B0 = h.bond_store()
    h.set_option("cg_reuse_forward", 1)
    h.bond_load(B)
    c1, _ = h.cgrad(4)
    B1 = h.bond_store()
    # class C (b=4) at this size is ill-conditioned: the third pass already amplifies 1e-16
    # differences to ~1e-6 (same effect between two float64 oracle orderings)
    tol = 1e-4 if b == 4 else 1e-10
    assert rel(c1[:2], c0[:2]) < 1e-10 and rel(c1, c0) < tol and rel(c1, costs_o) < max(tol, 1e-9)
    assert rel(B1, B0) < (1e-2 if b == 4 else 1e-6)
    with pytest.raises(capi.TnmlError):
        h.set_option("no_such_option", 1)
    h.close()


def test_fulltest_matches_oracle(capi):
    """inference (fulltest.cc / util.h fullTest) on a held-out set after a sweep."""
    feat, labels, W = make_problem(N=10, NT=900, m0=3)
    tr, te = slice(0, 600), slice(600, 900)
    h = _gpu_state(capi, feat[tr], labels[tr], W)
    p = capi.BondParams(3, 0.0, 1e-10, 1e-10, 8, 4, 0)
    for b, ha in O.sweep_schedule(10):
        h.bond_update(b, ha, p)
    Wt = h.get_mps()
    h.close()
    ncor_o, pred_o, P_o = O.full_test(Wt, feat[te], labels[te])
    from tnml_b200 import fixedl
    lines = []
    ncor, pred = fixedl.fullTest(Wt, feat[te], labels[te].astype(np.int32), log=lines.append)
    # near-ties of |P_l| may flip between the two summation orders
    assert abs(ncor - ncor_o) <= 2 and np.sum(pred != pred_o) <= 3
    assert lines[0].startswith(f"{ncor}/300 correct") and lines[-1] == "Total # test images = 300"
    assert ncor > 60     # learned something (chance = 30)


def test_svd_rank_deficient_isometry(capi):
    """ITensor's svd returns orthonormal U columns also for zero singular values; Minm can force such
    vectors to be kept (real MNIST border sites have phi = [1, 0] for every image, so the bond matrix
    is rank deficient there).  The device SVD completes the null-space columns: W(c) is an isometry,
    m / truncerr / U S V still match the oracle.  Direct Jacobi path (16 columns) and QR path (40)."""
    rng = np.random.default_rng(11)
    for m0, N, jc, cases in ((8, 12, 6, [(4, 1), (4, 2), (6, 1), (5, 2), (8, 1)]), (20, 16, 8, [(6, 1), (10, 2)])):
        feat, labels, W = make_problem(N=N, NT=64, m0=m0)
        for b, ha in cases:
            h = _gpu_state(capi, feat, labels, W)
            for bb in range(1, b):
                h.set_bond(bb)
                h.shift_env(bb, capi.FROMLEFT)
            h.set_bond(b)
            B0 = O.form_bond(W[b], W[b + 1])
            # rank-3 bond matrix whose odd rows and odd columns vanish exactly (like s = 1 slices on
            # border pixels), folded back into the tensor layout through bond_matrix of an index probe
            Im, _ = O.bond_matrix(np.arange(B0.size, dtype=np.float64).reshape(B0.shape), b, ha, jc)
            M = rng.standard_normal((Im.shape[0], 3)) @ rng.standard_normal((3, Im.shape[1]))
            M[1::2, :] = 0.0
            M[:, 1::2] = 0.0
            B = np.zeros(B0.size)
            B[Im.astype(np.int64).reshape(-1)] = M.reshape(-1)
            B = B.reshape(B0.shape)
            assert np.array_equal(O.bond_matrix(B, b, ha, jc)[0], M)
            keep = 6
            Wb, Wb1, m, te = O.svd_split(B, b, ha, jc, 8, keep, 1e-10)
            assert m == keep                                    # Minm forces 3 zero-sigma vectors in
            h.bond_load(B)
            gm, gte = h.svd_split(capi.FROMLEFT if ha == 1 else capi.FROMRIGHT, 1e-10, 8, keep)
            assert gm == m and abs(gte - te) <= 1e-9 * max(te, 1e-300) + 1e-24 * np.linalg.norm(B) ** 2
            gWb, gWb1 = h.get_site(b), h.get_site(b + 1)
            assert rel(O.form_bond(gWb, gWb1), B) < 1e-11, (m0, b, ha)
            iso = gWb if ha == 1 else gWb1
            if ha == 1:
                Ug = (np.transpose(iso, (0, 1, 3, 2)) if iso.ndim == 4 else iso).reshape(-1, m)
                assert np.abs(Ug.T @ Ug - np.eye(m)).max() < 1e-12, (m0, b, ha)
            else:
                Vg = iso.reshape(m, -1)
                assert np.abs(Vg @ Vg.T - np.eye(m)).max() < 1e-12, (m0, b, ha)
            h.close()
