"""Micro-benchmark of the device SVD (tnml_svd_split) through the C-ABI:
240x240 (class L bond, ml=mr=120) and 2400x240 (class C) matrices of several
spectra.  Run on the GPU box:  python tools/svd_bench.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnml_b200 import capi, data  # noqa: E402


def main(m=120, only=None):
    N, NT = 20, 64
    pix, labels = data.synthetic_digits(NT, 14, seed=1)
    feat = data.phi(pix[:, 80:80 + N])
    W = data.random_mps(N, 2, m, seed=2)
    h = capi.Handle(0)
    h.set_images(feat, labels)
    h.set_mps(W)
    h.init_envs()
    rng = np.random.default_rng(0)
    for b in (8, 9):
        for bb in range(1, b):
            h.set_bond(bb)
            h.shift_env(bb, capi.FROMLEFT)
        h.set_bond(b)
        h.bond_form()
        shape = h.bond_shape()
        n = int(np.prod(shape))
        rows = shape[0] * 2 * (10 if (len(shape) == 5 and b == 10) else 1)
        cols = n // rows
        cases = {}
        B0 = h.bond_store()
        cases["W(b)W(b+1) + 1e-3 noise"] = B0 + 1e-3 * np.linalg.norm(B0) / np.sqrt(n) * rng.standard_normal(shape)
        cases["random"] = rng.standard_normal(shape)
        U = rng.standard_normal((rows, min(rows, cols)))
        V = rng.standard_normal((min(rows, cols), cols))
        cases["graded 1e-9"] = ((U * np.logspace(0, -9, U.shape[1])) @ V).reshape(shape)
        for name, B in cases.items():
            if only and only not in name:
                continue
            for rep in range(2):
                h.bond_load(B)
                h.set_timing(True)
                h.stats(reset=True)
                t0 = time.perf_counter()
                mnew, te = h.svd_split(capi.FROMLEFT, 1e-10, m, m // 2)
                wall = time.perf_counter() - t0
                st = h.stats(reset=True)
                h.set_timing(False)
            Wb, Wb1 = h.get_site(b), h.get_site(b + 1)
            # gauge-free check against LAPACK
            M = B.reshape(rows, cols)
            s = np.linalg.svd(M, compute_uv=False)
            newB = (Wb.reshape(rows, mnew) @ Wb1.reshape(mnew, cols)) if len(shape) == 4 else None
            err = None
            if newB is not None:
                Uo, so, Vo = np.linalg.svd(M, full_matrices=False)
                ref = (Uo[:, :mnew] * so[:mnew]) @ Vo[:mnew]
                err = np.abs(newB - ref).max() / so[0]
            print(f"bond {b} {rows}x{cols} {name:28s} m={mnew} truncerr={te:.3e} (lapack {np.sum(s[mnew:]**2):.3e}) "
                  f"svd_ms={st.ms_svd:.3f} wall_ms={wall*1e3:.3f} launches={st.launches} newB_err={err}")
            # restore the sites for the next case
            h.set_site(b, W[b])
            h.set_site(b + 1, W[b + 1])
            h.set_bond(b + 1)
            h.set_bond(b)
    h.close()


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 120, sys.argv[2] if len(sys.argv) > 2 else None)
