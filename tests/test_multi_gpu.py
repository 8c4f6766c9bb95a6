"""Two ranks (one process per GPU, NCCL all-reduce of [gradient | 16 scalars]) against one rank holding
all images: costs, #correct, the first CG step and the CG cost lines agree to summation-order rounding,
and the two ranks hold bit-identical results (the SVD / CG vector algebra is replicated on them).
Needs 2 GPUs (`gpurun --gpus 2`); skipped otherwise."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


def _run(world, work):
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "_mg_worker.py"), str(r), str(world), str(work)])
             for r in range(world)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    return [np.load(os.path.join(work, f"rank{r}_of{world}.npz")) for r in range(world)]


def test_two_ranks_match_one_rank(tmp_path):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    (tmp_path / "w1").mkdir()
    (tmp_path / "w2").mkdir()
    one = _run(1, tmp_path / "w1")[0]
    two = _run(2, tmp_path / "w2")
    for k in one.files:
        if k == "bcast":
            continue
        assert np.array_equal(two[0][k], two[1][k]), k               # replicated state is bit-identical across ranks
    for bond in (3, 4):
        C1, C2 = float(one[f"C{bond}"]), float(two[0][f"C{bond}"])
        assert abs(C1 - C2) < 1e-12 * C1
        assert np.abs(one[f"CL{bond}"] - two[0][f"CL{bond}"]).max() < 1e-12 * C1
        assert int(one[f"nc{bond}"]) == int(two[0][f"nc{bond}"])
        B1, B2 = one[f"B{bond}"], two[0][f"B{bond}"]
        assert np.abs(B1 - B2).max() < 1e-11 * np.abs(B1).max()      # B + a p after one CG step
        assert abs(one[f"costs{bond}"][0] - two[0][f"costs{bond}"][0]) < 1e-11 * one[f"costs{bond}"][0]
    assert list(two[0]["bcast"]) == [3.5, 0.0] and list(two[1]["bcast"]) == [3.5, 0.0]   # tnml_comm_broadcast from rank 0
