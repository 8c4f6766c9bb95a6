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
