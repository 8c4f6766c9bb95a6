"""Checker for the tcgen05 int8 route (TEST INFRASTRUCTURE ONLY, like the rest of
oracle/): float64 matrix products through an Ozaki-style error-free splitting into signed 7-bit
slices with exact integer slice products.  First half: the scheme as it was sized in round 1
(`tools/ozaki_study.py`, truncated digits).  Second half (`slices_rn`, `oz_kernel_model`): the bit-level
model of the kernel that was built, which the GPU output is compared with bit for bit.

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


# ---- bit-level model of the tcgen05 kernel that was built (tnml_b200/csrc/tnml_ozaki.cu) -------------------
# The planned scheme above truncates digits (|q| <= 127); the kernel rounds them to nearest (|q| <= 64, scale to
# |x| <= 1/2), which halves the truncation error per plane.  The functions below restate the kernel's arithmetic
# operation by operation -- every step but the last two roundings is exact -- so a GPU test can ask for
# BIT-IDENTICAL output (tests/test_parity_shapes_gpu.py::test_tcgen05_env_advance_bit_exact).

_SHIFT = 6755399441055744.0   # 1.5 * 2^52


def slices_rn(M: np.ndarray, axis: int, s: int, shift_trick: bool = True):
    """`oz_slice_rows_kernel` (axis=1: one scale per row) / `oz_slice_cols_kernel` (axis=0: per column):
    2^e from frexp of the largest magnitude so that |x| 2^-e <= 1/2, then s round-to-nearest-even digits
    q = rint(128 x), x <- 128 x - q.  `shift_trick` computes the digit like `oz_digit` does (add and subtract
    1.5 * 2^52); False uses rint -- the two are bit-identical (tested).  Returns ([q_i] int64, 2^e)."""
    mx = np.max(np.abs(M), axis=axis, keepdims=True)
    _, e = np.frexp(mx)
    e = np.where(mx > 0, e + 1, 0)
    r = M * np.exp2(-e.astype(np.float64))          # exact (power of two)
    out = []
    for _ in range(s):
        r = r * 128.0                               # exact
        if shift_trick:
            t = r + _SHIFT                          # one rounding: to the nearest integer, ties to even
            q = t - _SHIFT                          # exact
        else:
            q = np.rint(r)
        out.append(q.astype(np.int64))
        r = r - q                                   # exact
    return out, np.exp2(e.astype(np.float64))


def oz_kernel_model(In: np.ndarray, Bm: np.ndarray, f1: np.ndarray, f2, S: int, ns: int = 8, div: int = 1) -> np.ndarray:
    """Out[row][j] = sum_p w_p(row) sum_a In[row][a] Bm[(a*S+p)][j] exactly as `oz_gemm_kernel<ns,S>` rounds it.
    In [rows][K], Bm [K*S][J] (row index a*S+p), f1/f2 [images][2] features, image of a row = row // div.
    Steps: planes of In per row and of Bm per column c = (j, p); exact integer level sums acc_L = sum_{i+j=L} A_i B_j^T;
    exact merge hi = ((acc0*128+acc1)*16384 + acc2*128+acc3), lo likewise from levels 4..7; v = RN(hi 2^(f-35) + lo 2^(f-63))
    (one fused multiply-add on exact operands = one rounding); o = fma(w_p, v_p, o) for p = 0..S-1, w_p = weight * 2^e(row).
    The fused multiply-adds of the last step are evaluated with exact rational arithmetic and rounded once."""
    from fractions import Fraction
    rows, K = In.shape
    J = Bm.shape[1]
    assert Bm.shape[0] == K * S and ns <= 8
    A, ea = slices_rn(In, 1, ns)
    cols = Bm.reshape(K, S, J)                                     # [a][p][j]
    Bq, eb = slices_rn(cols.reshape(K, S * J), 0, ns)              # planes [K][(p, j)], scale per column
    acc = [np.zeros((rows, S * J), dtype=np.int64) for _ in range(8)]
    for i in range(ns):
        for j in range(ns - i):
            acc[i + j] += A[i] @ Bq[j]
    p01, p23 = acc[0] * 128 + acc[1], acc[2] * 128 + acc[3]
    p45, p67 = acc[4] * 128 + acc[5], acc[6] * 128 + acc[7]
    hi, lo = p01 * 16384 + p23, p45 * 16384 + p67
    assert np.abs(hi).max(initial=0) < 2 ** 51 and np.abs(lo).max(initial=0) < 2 ** 51
    v = (hi.astype(np.float64) * (eb * 2.0 ** -35)) + (lo.astype(np.float64) * (eb * 2.0 ** -63))   # both products exact
    v = v.reshape(rows, S, J)
    img = np.arange(rows) // div
    if S == 2:
        w = np.stack([f1[img, 0], f1[img, 1]], axis=1)
    else:
        w = np.stack([f1[img, 0] * f2[img, 0], f1[img, 0] * f2[img, 1], f1[img, 1] * f2[img, 0], f1[img, 1] * f2[img, 1]], axis=1)
    w = w * ea                                                     # exact: ea is a power of two
    out = np.zeros((rows, J))
    for r in range(rows):
        wf = [Fraction(float(x)) for x in w[r]]
        for j in range(J):
            o = 0.0
            for p in range(S):
                o = float(wf[p] * Fraction(float(v[r, p, j])) + Fraction(o))   # fma: exact, then one rounding to nearest even
            out[r, j] = o
    return out


def fat_forward_model(Q: np.ndarray, F: np.ndarray) -> np.ndarray:
    """P[n][l] = sum_f Q[n][f] F[n][l][f] exactly as `fat_kernel_t<mode, MCH>` (tnml_kernels.cu, m <= 128) rounds it:
    lane i of the image's warp chains fma over f = i, i + 32, i + 64, i + 96 (starting from 0), then the 32 partial
    sums are combined by the xor butterfly 16, 8, 4, 2, 1 (plain float64 adds; every lane ends with the same bits)."""
    from fractions import Fraction
    n_img, m = Q.shape
    NL = F.shape[1]
    assert m <= 128 and F.shape == (n_img, NL, m)
    part = np.zeros((n_img, NL, 32))
    for n in range(n_img):
        qf = [Fraction(float(x)) for x in Q[n]]
        for l in range(NL):
            row = F[n, l]
            for lane in range(32):
                o = 0.0
                for f in range(lane, m, 32):
                    o = float(qf[f] * Fraction(float(row[f])) + Fraction(o))
                part[n, l, lane] = o
    idx = np.arange(32)
    for o in (16, 8, 4, 2, 1):
        part = part + part[:, :, idx ^ o]
    return part[:, :, 0]
