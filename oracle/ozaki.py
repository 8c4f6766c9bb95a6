"""Checker for the planned tcgen05 int8 route (TEST INFRASTRUCTURE ONLY, like the rest of
oracle/): float64 matrix products through an Ozaki-style error-free splitting into signed 7-bit
slices with exact integer slice products.  `tools/ozaki_study.py` sizes the number of slices with
it; a future GPU kernel is compared against `ozaki_matmul` bit for bit (the integer part is exact,
only the final float64 combination order matters).

No reference file:line -- the reference computes these contractions in float64 through ITensor
(`fixedL.cc:377,379,399,416,418`); this module restates the arithmetic a sliced kernel performs.
"""
import numpy as np


def slices(M: np.ndarray, axis: int, s: int):
    """Scale along `axis` (1: one power of two per row, 0: per column) to |x| < 1 and cut into
    s signed 7-bit integers: M ~ 2^e * sum_i 2^(-7(i+1)) q_i,  |q_i| <= 127.  Returns ([q_i], 2^e)."""
    mx = np.max(np.abs(M), axis=axis, keepdims=True)
    e = np.where(mx > 0, np.ceil(np.log2(np.where(mx > 0, mx, 1.0))) + 1, 0.0)
    r = M / np.exp2(e)
    out = []
    for _ in range(s):
        r = r * 128.0
        q = np.trunc(r)
        out.append(q.astype(np.int64))
        r = r - q
    return out, np.exp2(e)


def ozaki_matmul(A: np.ndarray, B: np.ndarray, s: int) -> np.ndarray:
    """A @ B with both operands cut into s slices; the s(s+1)/2 slice products with i + j < s are
    exact integers, summed per level i + j, levels combined in float64 from the smallest up."""
    As, ea = slices(A, 1, s)
    Bs, eb = slices(B, 0, s)
    C = np.zeros((A.shape[0], B.shape[1]))
    for lev in range(s - 1, -1, -1):
        acc = np.zeros((A.shape[0], B.shape[1]), dtype=np.int64)
        for i in range(lev + 1):
            acc += As[i] @ Bs[lev - i]
        C += acc.astype(np.float64) * 2.0 ** (-7 * (lev + 2))
    return C * ea * eb
