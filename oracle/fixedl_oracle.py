"""CPU oracle for TNML's `fixedL` per-bond hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a float64 numpy *restatement* of the reference algorithm
(/root/reference/fixedL.cc, paralleldo.h, util.h, mllib/mnist.h, image.h).
It is the checker that the CUDA path is compared with.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it; the product (`tnml_b200/`) never does.

PARITY UNPINNED: the reference's arithmetic lives in ITensor v2 (un-vendored,
unpinned: Makefile.sample:1,4-5) which is absent from /root/reference and from
this image, and the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4 / 8c).  The oracle is therefore pinned only by
  (1) two independent formulations agreeing -- `literal` (dense t.v exactly as
      fixedL.cc:183-185 builds it) vs `structured` (Khatri-Rao form),
  (2) mathematical invariants (finite-difference gradient, full-contraction
      `toverlap` == environment recursion, SVD reconstruction/orthogonality),
  (3) known-answer facts about the MNIST files (md5s, label histogram, the
      per-label selection order of readMNIST).
ITensor semantics that could not be read offline (SVD truncation rule,
DoRelCutoff default) are flagged `ASSUMED` below.

Conventions: sites are 1-indexed (1..N) like the reference.  A site tensor is
an ndarray [ml, d, mr] or, for the label site jc = N//2 (fixedL.cc:616),
[ml, d, mr, NL].  A bond tensor is [ml, d, d, mr] or [ml, d, d, mr, NL].
Environment slots: slot[j] is the left env through site j (dim = link(j,j+1))
or the right env from site j (dim = link(j-1,j)); a label-carrying env is
stored [NT, NL, m] ("fat"), a label-free one [NT, m] ("thin").
"""
from __future__ import annotations

import gzip
import os
import struct

import numpy as np

NL = 10  # fixedL.cc:15


# --------------------------------------------------------------------------
# data path: mllib/mnist.h, image.h, fixedL.cc:637-653
# --------------------------------------------------------------------------
def read_idx(path: str) -> np.ndarray:
    """idx1/idx3 reader (mllib/mnist.h:157-227: big-endian magic, counts, raw u8)."""
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    magic = struct.unpack(">I", raw[:4])[0]
    if magic == 0x803:  # 2051 images (mnist.h:167)
        n, r, c = struct.unpack(">III", raw[4:16])
        return np.frombuffer(raw, np.uint8, n * r * c, 16).reshape(n, r * c)
    if magic == 0x801:  # 2049 labels (mnist.h:205)
        n = struct.unpack(">I", raw[4:8])[0]
        return np.frombuffer(raw, np.uint8, n, 8)
    raise ValueError(f"bad idx magic {magic:#x} in {path}")


def read_mnist(datadir: str, kind: str = "Train", NT: int = 50000):
    """readMNIST (mllib/mnist.h:443-530): first NT images *per label* in file
    order, pixels divided by 255 (mnist.h:495).  Returns (data[n,784] f64 in
    [0,1], labels[n] int, file_index[n])."""
    pre = "train" if kind == "Train" else "t10k"
    imgs = read_idx(os.path.join(datadir, f"{pre}-images-idx3-ubyte"))
    labs = read_idx(os.path.join(datadir, f"{pre}-labels-idx1-ubyte"))
    return select_per_label(imgs, labs, NT)


def select_per_label(imgs_u8: np.ndarray, labs: np.ndarray, NT: int):
    """The per-label cap loop of mnist.h:472-496."""
    counts = np.zeros(NL, np.int64)
    keep = []
    for i, l in enumerate(labs):
        if counts[l] >= NT:
            continue
        counts[l] += 1
        keep.append(i)
    keep = np.asarray(keep, np.int64)
    data = imgs_u8[keep].astype(np.float64) / 255.0
    return data, labs[keep].astype(np.int64), keep


def reduce_image(data: np.ndarray, newlen: int) -> np.ndarray:
    """image.h:316-346 `reduce`: block mean, kept real (no rounding).
    data[n, L*L] raster order pixel(x,y) at y*L+x  ->  [n, newlen*newlen].
    fixedL.cc never calls this; `imglen` is our documented add-on."""
    n, npix = data.shape
    L = int(round(np.sqrt(npix)))
    if newlen == L:
        return data.copy()
    bsize = L // newlen
    rem = L % bsize
    img = data.reshape(n, L, L)  # [n, y, x]
    img = img[:, rem:rem + bsize * newlen, rem:rem + bsize * newlen]
    out = img.reshape(n, newlen, bsize, newlen, bsize).sum(axis=(2, 4)) / (bsize * bsize)
    return out.reshape(n, newlen * newlen)


def phi(g: np.ndarray, d: int = 2) -> np.ndarray:
    """Local feature map fixedL.cc:637-642: x = g/255 (a SECOND /255),
    phi_n = (x/4)^(n-1).  g is the already-/255 pixel.  Returns [..., d]."""
    g = np.asarray(g, np.float64)
    if np.any(g < 0) or np.any(g > 255.0):
        raise ValueError("Expected g to be in [0,255]")  # fixedL.cc:639
    x = g / 255.0
    return np.stack([np.power(x / 4.0, n) for n in range(d)], axis=-1)


def features(data: np.ndarray, d: int = 2) -> np.ndarray:
    """TState::data (fixedL.cc:39-46): feat[n, j-1, k-1] = phi(img(j), k)."""
    return phi(data, d)


def shard_bounds(nshard: int, ntask: int):
    """ParallelDo(Nthread,Ntask) (paralleldo.h:32-43): contiguous ranges of
    Ntask//Nthread, the last absorbs the remainder."""
    th = ntask // nshard
    b = [(n * th, (n + 1) * th) for n in range(nshard)]
    b[-1] = (b[-1][0], ntask)
    return b


# --------------------------------------------------------------------------
# deterministic initial W (replaces the time-seeded Global::random())
# --------------------------------------------------------------------------
def link_dims(N: int, d: int, m: int):
    """link(j) between sites j and j+1, j=0..N (0 and N are dummy dim-1 links)."""
    dims = [1] * (N + 1)
    for j in range(1, N):
        cap = min(j, N - j)
        dims[j] = m if cap >= 40 else min(m, d ** cap)
    return dims


def random_mps(N: int, d: int = 2, m: int = 10, seed: int = 1, jc: int | None = None,
               noise: float = 0.3):
    """Seeded start MPS, orthogonality centre at site 1 (like ITensor `sum`,
    SURVEY 8c(5)), label index on site jc, W[jc] /= norm (fixedL.cc:725).
    s=0 slices are near-identity so environments stay O(1) for phi=[1,~1e-3]."""
    jc = N // 2 if jc is None else jc
    rng = np.random.default_rng(seed)
    dims = link_dims(N, d, m)
    W = [None] * (N + 1)
    for j in range(1, N + 1):
        ml, mr = dims[j - 1], dims[j]
        shape = (ml, d, mr, NL) if j == jc else (ml, d, mr)
        A = noise * rng.standard_normal(shape) / np.sqrt(max(ml, mr))
        eye = np.eye(ml, mr)
        if j == jc:
            A[:, 0, :, :] += eye[:, :, None] * (1.0 + 0.5 * rng.standard_normal(NL))[None, None, :]
        else:
            A[:, 0, :] += eye
        W[j] = A
    # right-canonicalise N..2 (centre ends on site 1)
    for j in range(N, 1, -1):
        A = W[j]
        ml = A.shape[0]
        Q, R = np.linalg.qr(A.reshape(ml, -1).T)  # A^T = Q R -> A = R^T Q^T
        k = Q.shape[1]
        W[j] = Q.T.reshape((k,) + A.shape[1:])
        Wp = W[j - 1]  # contract R^T into the right link (axis 2) of site j-1
        R = R / np.linalg.norm(R) * np.sqrt(k)   # keep magnitudes O(1) along the chain
        W[j - 1] = np.moveaxis(np.tensordot(Wp, R.T, axes=([2], [0])), -1, 2)
    W[1] = W[1] / np.linalg.norm(W[1])
    W[jc] = W[jc] / np.linalg.norm(W[jc])
    return W


def toverlap(W, feat_n: np.ndarray, jc: int):
    """util.h:19-40: full contraction of one image with W -> label vector."""
    N = len(W) - 1
    r = None
    for j in range(N, jc, -1):
        M = np.tensordot(feat_n[j - 1], W[j], axes=([0], [1]))  # [ml, mr]
        r = M[:, 0] if r is None else M @ r
    l = None
    for j in range(1, jc):
        M = np.tensordot(feat_n[j - 1], W[j], axes=([0], [1]))
        l = M[0, :] if l is None else l @ M
    Mc = np.tensordot(feat_n[jc - 1], W[jc], axes=([0], [1]))  # [ml, mr, NL]
    if l is None:
        l = np.ones(1)
    if r is None:
        r = np.ones(1)
    return np.einsum("a,abl,b->l", l, Mc, r)


def full_test(W, feat: np.ndarray, labels: np.ndarray):
    """util.h:123-200 fullTest: per image W_l = toverlap(psi,img,cent), prediction
    argmax_l |W_l| (first strict maximum).  Returns (ncorrect, predictions, outputs)."""
    N = len(W) - 1
    jc = [j for j in range(1, N + 1) if W[j].ndim == 4][0]
    P = np.stack([toverlap(W, feat[n], jc) for n in range(feat.shape[0])])
    pred = argmax_first(np.abs(P))
    return int(np.sum(pred == np.asarray(labels))), pred, P


# --------------------------------------------------------------------------
# TrainStates (fixedL.cc:64-274), structured form
# --------------------------------------------------------------------------
class TrainStates:
    def __init__(self, feat: np.ndarray, labels: np.ndarray, nshard: int = 1):
        self.feat = np.ascontiguousarray(feat, np.float64)  # [NT, N, d]
        self.labels = np.asarray(labels, np.int64)
        self.NT, self.N, self.d = self.feat.shape
        self.jc = self.N // 2
        self.slot = [None] * (self.N + 2)
        self.currb = -1
        self.bounds = shard_bounds(nshard, self.NT)

    def size(self):
        return self.NT

    # --- site "transfer" of one env through site j ------------------------
    def _site_mats(self, Wj, j):
        """A_n(j)*W(j): [NT, ml, mr(,NL)]"""
        return np.tensordot(self.feat[:, j - 1, :], Wj, axes=([1], [1]))

    def init(self, W):
        """fixedL.cc:122-157: right envs E_N..E_3 then setBond(1)."""
        N = self.N
        for n in range(N, 2, -1):
            self.slot[n] = self._advance(None if n == N else self.slot[n + 1], W[n], n, "right")
        self.currb = -1
        self.set_bond(1)

    def _advance(self, prev, Wc, c, side):
        """One env step (fixedL.cc:144-150, 221-229).  side='left': new left
        env through c from left env through c-1; 'right': mirror."""
        M = self._site_mats(Wc, c)  # [NT, ml, mr] or [NT, ml, mr, NL]
        has_lab = (M.ndim == 4)
        if side == "left":
            if prev is None:
                out = M[:, 0]                         # [NT, mr] or [NT, mr, NL]
                return np.moveaxis(out, -1, 1) if has_lab else out
            if prev.ndim == 3:                         # fat prev [NT, NL, ml]
                return np.einsum("nla,nab->nlb", prev, M)
            if has_lab:
                return np.einsum("na,nabl->nlb", prev, M)
            return np.einsum("na,nab->nb", prev, M)
        else:
            if prev is None:
                out = M[:, :, 0]                      # [NT, ml] or [NT, ml, NL]
                return np.moveaxis(out, -1, 1) if has_lab else out
            if prev.ndim == 3:
                return np.einsum("nab,nlb->nla", M, prev)
            if has_lab:
                return np.einsum("nabl,nb->nla", M, prev)
            return np.einsum("nab,nb->na", M, prev)

    def set_bond(self, b):
        """fixedL.cc:159-190.  Structured form: only selects the env pair."""
        if self.currb == b:
            return
        self.currb = b

    def envs(self, b=None):
        b = self.currb if b is None else b
        LE = self.slot[b - 1] if b - 1 > 0 else None
        RE = self.slot[b + 2] if b + 2 < self.N + 1 else None
        return LE, RE

    def shiftE(self, W, b, direction):
        """fixedL.cc:192-233."""
        if direction == "Fromleft":
            c, prevc = b, b - 1
            prev = self.slot[prevc] if prevc >= 1 else None
            self.slot[c] = self._advance(prev, W[c], c, "left")
        else:
            c, prevc = b + 1, b + 2
            prev = self.slot[prevc] if prevc <= self.N else None
            self.slot[c] = self._advance(prev, W[c], c, "right")

    # --- projected input, dense (literal mode, fixedL.cc:183-185) ----------
    def dense_v(self, b, sl=slice(None)):
        """t.v = A(b) x A(b+1) x LE x RE as [n, ml, d, d, mr, (NL)]."""
        LE, RE = self.envs(b)
        x = self.feat[sl, b - 1]
        y = self.feat[sl, b]
        n = x.shape[0]
        l = np.ones((n, 1)) if LE is None else LE[sl]
        r = np.ones((n, 1)) if RE is None else RE[sl]
        if l.ndim == 3:   # fat left [n, NL, ml]
            return np.einsum("nla,ns,nt,nb->nastbl", l, x, y, r)
        if r.ndim == 3:
            return np.einsum("na,ns,nt,nlb->nastbl", l, x, y, r)
        return np.einsum("na,ns,nt,nb->nastb", l, x, y, r)


def _lr(ts: TrainStates, sl=slice(None)):
    LE, RE = ts.envs()
    b = ts.currb
    x = ts.feat[sl, b - 1]
    y = ts.feat[sl, b]
    n = x.shape[0]
    l = np.ones((n, 1)) if LE is None else LE[sl]
    r = np.ones((n, 1)) if RE is None else RE[sl]
    return l, x, y, r


def project(B, ts: TrainStates, sl=slice(None), literal=False):
    """P_n[l] = B * t.v  (fixedL.cc:318, 377, 399, 416)."""
    if literal:
        v = ts.dense_v(ts.currb, sl)
        if B.ndim == 5 and v.ndim == 5:
            return np.einsum("astbl,nastb->nl", B, v)
        return np.einsum("astb,nastbl->nl", B, v)
    l, x, y, r = _lr(ts, sl)
    if B.ndim == 5:                       # class C: label on the bond tensor
        n, ml = l.shape
        d, mr = B.shape[2], B.shape[3]
        LX = (l[:, :, None] * x[:, None, :]).reshape(n, -1)                 # [n, (a s)]
        T = (LX @ B.reshape(LX.shape[1], -1)).reshape(n, d, mr, NL)          # BLAS: [n, t, b, l]
        YR = y[:, :, None] * r[:, None, :]                                   # [n, t, b]
        return np.einsum("ntbl,ntb->nl", T, YR)
    n = x.shape[0]
    ml, d, _, mr = B.shape
    if r.ndim == 3:                       # class L: fat right env
        LX = (l[:, :, None] * x[:, None, :]).reshape(n, -1)                 # [n, (a s)]
        T = (LX @ B.reshape(ml * d, d * mr)).reshape(n, d, mr)               # BLAS: [n, t, b]
        Q = np.einsum("ntb,nt->nb", T, y)
        return np.einsum("nb,nlb->nl", Q, r)
    YR = (y[:, :, None] * r[:, None, :]).reshape(n, -1)                     # class R: [n, (t b)]
    T = (YR @ B.reshape(ml * d, d * mr).T).reshape(n, ml, d)                 # BLAS: [n, a, s]
    Q = np.einsum("nas,ns->na", T, x)
    return np.einsum("na,nla->nl", Q, l)


def backproject(dP, Bshape, ts: TrainStates, sl=slice(None), literal=False):
    """sum_n dP_n * dag(t.v_n)  (fixedL.cc:379, 418) -> tensor shaped like B."""
    if literal:
        v = ts.dense_v(ts.currb, sl)
        if len(Bshape) == 5:
            return np.einsum("nl,nastb->astbl", dP, v)
        return np.einsum("nl,nastbl->astb", dP, v)
    l, x, y, r = _lr(ts, sl)
    if len(Bshape) == 5:
        n = l.shape[0]
        LX = (l[:, :, None] * x[:, None, :]).reshape(n, -1)                 # [n, (a s)]
        YRP = ((y[:, :, None] * r[:, None, :])[:, :, :, None] * dP[:, None, None, :]).reshape(n, -1)
        return (LX.T @ YRP).reshape(Bshape)                                  # BLAS: [(a s), (t b l)]
    n = x.shape[0]
    if r.ndim == 3:
        Z = np.einsum("nl,nlb->nb", dP, r)
        LX = (l[:, :, None] * x[:, None, :]).reshape(n, -1)                 # [n, (a s)]
        YZ = (y[:, :, None] * Z[:, None, :]).reshape(n, -1)                 # [n, (t b)]
        return (LX.T @ YZ).reshape(Bshape)                                   # BLAS
    Z = np.einsum("nl,nla->na", dP, l)
    ZX = (Z[:, :, None] * x[:, None, :]).reshape(n, -1)
    YR = (y[:, :, None] * r[:, None, :]).reshape(n, -1)
    return (ZX.T @ YR).reshape(Bshape)


def argmax_first(w: np.ndarray) -> np.ndarray:
    """util.h:42-57: first strict maximum."""
    return np.argmax(w, axis=-1)  # numpy argmax returns the first maximum


def quadcost(B, ts: TrainStates, lam: float = 0.0, literal=False, detail=False):
    """fixedL.cc:280-344.  Returns un-normalised C (and, with detail, the
    per-label costs and the number correct)."""
    CL = np.zeros(NL)
    ncor = 0
    for (b0, b1) in ts.bounds:      # per-"thread" partials, reduced in order
        sl = slice(b0, b1)
        P = project(B, ts, sl, literal)
        lab = ts.labels[sl]
        dP = -P
        dP[np.arange(len(lab)), lab] += 1.0
        e = np.sum(dP * dP, axis=1)
        CL += np.bincount(lab, weights=e, minlength=NL)
        ncor += int(np.sum(argmax_first(np.abs(P)) == lab))
    C = float(np.sum(CL)) + lam * float(np.sum(B * B))
    if detail:
        return C, CL, ncor
    return C


def _grad(B, ts, lam, literal):
    G = np.zeros_like(B)
    C = 0.0
    for (b0, b1) in ts.bounds:
        sl = slice(b0, b1)
        P = project(B, ts, sl, literal)
        lab = ts.labels[sl]
        dP = -P
        dP[np.arange(len(lab)), lab] += 1.0
        G += backproject(dP, B.shape, ts, sl, literal)
        C += float(np.sum(dP * dP))
    if lam != 0.0:
        G = G - lam * B
    return G, C


def cgrad(B, ts: TrainStates, Npass: int = 4, lam: float = 0.0, cconv: float = 1e-10,
          literal=False, log=None):
    """fixedL.cc:349-445, operation for operation."""
    NT = ts.size()
    B = B.copy()
    r, _ = _grad(B, ts, lam, literal)                    # 373-386
    p = r.copy()                                         # 388
    costs, rnorms = [], []
    for ps in range(1, Npass + 1):                       # 389
        pAp = 0.0                                        # 393-403
        for (b0, b1) in ts.bounds:
            pv = project(p, ts, slice(b0, b1), literal)
            pAp += float(np.sum(pv * pv))
        pAp += lam * float(np.sum(p * p))
        a = float(np.sum(r * r)) / pAp                   # 405
        B = B + a * p                                    # 406
        if ps == Npass:                                  # 409
            break
        nr, C = _grad(B, ts, lam, literal)               # 412-422
        beta = float(np.sum(nr * nr)) / float(np.sum(r * r))  # 423
        r = nr
        C += lam * float(np.sum(B * B))                  # 427-428
        costs.append(C / NT)                             # 429 prints C/NT
        rn = float(np.sqrt(np.sum(r * r)))
        rnorms.append(rn)
        if log:
            log(f"  Cost = {C / NT:.10f}")
        if rn < cconv:                                   # 432-436
            break
        p = r + beta * p                                 # 442
    return B, costs, rnorms


# --------------------------------------------------------------------------
# svd + truncation (ITensor v2 svd(B,U,S,V,{Cutoff,Maxm,Minm}), fixedL.cc:519-521)
# --------------------------------------------------------------------------
def truncate_spectrum(P: np.ndarray, maxm: int, minm: int, cutoff: float,
                      do_rel_cutoff: bool = False):
    """ASSUMED (recalled from ITensor v2 svdalgs `truncate`):
    P = sigma^2 descending.  Drop from the tail while count > maxm; then keep
    dropping while truncerr + P_n < cutoff*scale and count > minm.
    Returns (m_kept, truncerr)."""
    P = np.asarray(P, np.float64)
    m = len(P)
    truncerr = 0.0
    while m > maxm:
        truncerr += P[m - 1]
        m -= 1
    scale = 1.0
    if do_rel_cutoff:
        scale = float(np.sum(P))
        if scale == 0.0:
            scale = 1.0
    while m > minm and m > 1 and truncerr + P[m - 1] < cutoff * scale:
        truncerr += P[m - 1]
        m -= 1
    return m, truncerr / scale


def bond_matrix(B, b, ha, jc):
    """Rows = indices B shares with the OLD W.A(c).  Returns the
    matrix and the tensor shapes needed to fold U / S*V back into sites."""
    has_lab = (B.ndim == 5)
    ml, d, _, mr = B.shape[:4]
    lab_on_b = has_lab and (b == jc)        # label sits on site b
    lab_on_b1 = has_lab and (b + 1 == jc)   # label sits on site b+1
    if ha == 1:   # c = b : rows (alpha, s [,L]), cols (t, beta [,L])
        if lab_on_b:
            M = np.transpose(B, (0, 1, 4, 2, 3)).reshape(ml * d * NL, d * mr)
        elif lab_on_b1:
            M = B.reshape(ml * d, d * mr * NL)
        else:
            M = B.reshape(ml * d, d * mr)
    else:         # c = b+1 : rows (t, beta [,L]), cols (alpha, s [,L])
        if lab_on_b1:
            M = np.transpose(B, (2, 3, 4, 0, 1)).reshape(d * mr * NL, ml * d)
        elif lab_on_b:
            M = np.transpose(B, (2, 3, 0, 1, 4)).reshape(d * mr, ml * d * NL)
        else:
            M = np.transpose(B, (2, 3, 0, 1)).reshape(d * mr, ml * d)
    return M, (ml, d, mr, lab_on_b, lab_on_b1)


def svd_split(B, b, ha, jc, maxm, minm, cutoff, do_rel_cutoff=False):
    """fixedL.cc:519-521: W(c) <- U, W(c+dc) <- S*V.  Returns
    (W_b, W_b1, newm, truncerr)."""
    M, (ml, d, mr, lab_b, lab_b1) = bond_matrix(B, b, ha, jc)
    U, s, Vt = np.linalg.svd(M, full_matrices=False)
    m, terr = truncate_spectrum(s * s, maxm, minm, cutoff, do_rel_cutoff)
    U = U[:, :m]
    SV = s[:m, None] * Vt[:m]
    if ha == 1:
        Wb = (np.transpose(U.reshape(ml, d, NL, m), (0, 1, 3, 2)) if lab_b
              else U.reshape(ml, d, m))
        Wb1 = SV.reshape(m, d, mr, NL) if lab_b1 else SV.reshape(m, d, mr)
    else:
        Wb1 = U.T.reshape(m, d, mr, NL) if lab_b1 else U.T.reshape(m, d, mr)
        Wb = (np.transpose(SV.reshape(m, ml, d, NL), (1, 2, 0, 3)) if lab_b
              else np.transpose(SV.reshape(m, ml, d), (1, 2, 0)))
    return Wb, Wb1, m, terr


def form_bond(Wb, Wb1):
    """oB = W.A(c)*W.A(c+dc) (fixedL.cc:494) as [ml,d,d,mr(,NL)]."""
    if Wb.ndim == 4:
        return np.einsum("asml,mtb->astbl", Wb, Wb1)
    if Wb1.ndim == 4:
        return np.einsum("asm,mtbl->astbl", Wb, Wb1)
    return np.einsum("asm,mtb->astb", Wb, Wb1)


def sweep_schedule(N):
    """sweepnext: b=1..N-1 with ha=1, then b=N-1..1 with ha=2."""
    return [(b, 1) for b in range(1, N)] + [(b, 2) for b in range(N - 1, 0, -1)]


def mldmrg(W, ts: TrainStates, Nsweep, maxm, minm, cutoff, Npass=4, lam=0.0, cconv=1e-10,
           literal=False, do_rel_cutoff=False, log=None, max_bonds=None, record=None):
    """fixedL.cc:451-570.  Mutates W and ts; returns a list of per-bond dicts
    (cost after SVD / NT, ncorrect, new m, truncerr, CG costs)."""
    N, NT, jc = ts.N, ts.NT, ts.jc
    out = []
    nb = 0
    for sw in range(1, Nsweep + 1):
        for (b, ha) in sweep_schedule(N):
            c, dc = (b, +1) if ha == 1 else (b + 1, -1)
            ts.set_bond(b)                                     # 488
            if log:
                log(f"Sweep {sw} Half {ha} Bond {c}")
            oB = form_bond(W[b], W[b + 1])                     # 493-498
            origm = W[b].shape[2]
            B, costs, rnorms = cgrad(oB, ts, Npass, lam, cconv, literal, log)  # 504
            Wb, Wb1, newm, terr = svd_split(B, b, ha, jc, maxm, minm, cutoff, do_rel_cutoff)
            W[b], W[b + 1] = Wb, Wb1                           # 519-521
            newB = form_bond(W[b], W[b + 1])                   # 527
            C, CLab, ncor = quadcost(newB, ts, lam, literal, detail=True)  # 532
            if log:
                log(f"SVD trunc err = {terr:.2E}")
                log(f"Original m={origm}, New m={newm}")
                log(f"--> After SVD, Cost = {C / NT:.10f}")
            ts.shiftE(W, b, "Fromleft" if ha == 1 else "Fromright")       # 540
            rec = dict(sweep=sw, half=ha, b=b, c=c, cost=C / NT, ncor=ncor, m=newm,
                       truncerr=terr, cg_costs=costs, cg_rnorms=rnorms,
                       dB=float(np.linalg.norm(B - newB)), Bnorm=float(np.linalg.norm(B)))
            if record is not None:
                rec.update(record(B, newB))
            out.append(rec)
            nb += 1
            if max_bonds is not None and nb >= max_bonds:
                return out
    return out


# --------------------------------------------------------------------------
# synthetic MNIST-shaped data (bench / tests on the GPU box, where
# /root/reference does not exist): SURVEY 8d config 5 recipe
# --------------------------------------------------------------------------
def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def synthetic_pixels(NT: int, npix: int, seed: int = 20260925, first: int = 0):
    """u8 pixels: zero with prob 0.81 else uniform{1..255}; labels = image % 10.
    Counter-based, so shard `first:first+NT` of a larger set is reproducible."""
    with np.errstate(over="ignore"):
        idx = (np.arange(first, first + NT, dtype=np.uint64)[:, None] * np.uint64(npix)
               + np.arange(npix, dtype=np.uint64)[None, :])
        h = _splitmix64(idx ^ _splitmix64(np.full(1, seed, np.uint64)))
        u = (h >> np.uint64(40)).astype(np.float64) / float(1 << 24)
        v = (_splitmix64(h) >> np.uint64(40)).astype(np.float64) / float(1 << 24)
    pix = np.where(u < 0.81, 0, 1 + np.floor(v * 255)).astype(np.uint8)
    labels = (np.arange(first, first + NT) % NL).astype(np.int64)
    return pix, labels


# --------------------------------------------------------------------------
# initial W (fixedL.cc:682-728, util.h:76-121) -- checker for
# tnml_b200/host/initial_w.h.  ITensor's sum(vector<MPS>,args) ASSUMED as in
# pairwise direct sums, each followed by orthogonalize(args)
# (left half sweep without truncation, right half sweep truncating with the
# SVD rule of truncate_spectrum).  Raw MPS = list (1-indexed) of [ml, p, mr].
# --------------------------------------------------------------------------
def mps_product_state(feat_n: np.ndarray):
    """util.h:76-102 makeMPS: feat_n [N, d] -> bond-dimension-1 MPS."""
    return [None] + [feat_n[j].reshape(1, -1, 1).copy() for j in range(feat_n.shape[0])]


def mps_direct_sum(A, B):
    N = len(A) - 1
    C = [None]
    for j in range(1, N + 1):
        x, y = A[j], B[j]
        ml = 1 if j == 1 else x.shape[0] + y.shape[0]
        mr = 1 if j == N else x.shape[2] + y.shape[2]
        z = np.zeros((ml, x.shape[1], mr))
        lo = 0 if j == 1 else x.shape[0]
        ro = 0 if j == N else x.shape[2]
        z[:x.shape[0], :, :x.shape[2]] += x
        z[lo:lo + y.shape[0], :, ro:ro + y.shape[2]] += y
        C.append(z)
    return C


def mps_orthogonalize(W, cutoff, maxm, do_rel_cutoff=False):
    N = len(W) - 1
    for j in range(N, 1, -1):
        ml, p, mr = W[j].shape
        U, s, Vt = np.linalg.svd(W[j].reshape(ml, p * mr), full_matrices=False)
        k = max(1, int(np.sum(s > 1e-14 * s[0])))
        W[j] = Vt[:k].reshape(k, p, mr)
        W[j - 1] = np.tensordot(W[j - 1], U[:, :k] * s[:k], axes=([2], [0]))
    for j in range(1, N):
        ml, p, mr = W[j].shape
        U, s, Vt = np.linalg.svd(W[j].reshape(ml * p, mr), full_matrices=False)
        m, _ = truncate_spectrum(s * s, maxm, 1, cutoff, do_rel_cutoff)
        W[j] = U[:, :m].reshape(ml, p, m)
        W[j + 1] = np.tensordot(s[:m, None] * Vt[:m], W[j + 1], axes=([1], [0]))
    return W


def mps_sum(terms, cutoff, maxm, do_rel_cutoff=False):
    if len(terms) == 1:
        return terms[0]
    if len(terms) == 2:
        return mps_orthogonalize(mps_direct_sum(terms[0], terms[1]), cutoff, maxm, do_rel_cutoff)
    nt = [mps_orthogonalize(mps_direct_sum(terms[n], terms[n + 1]), cutoff, maxm, do_rel_cutoff)
          for n in range(0, len(terms) - 1, 2)]
    if len(terms) % 2 == 1:
        nt.append(terms[-1])
    return mps_sum(nt, cutoff, maxm, do_rel_cutoff)


def mps_overlap(A, B):
    E = np.ones((1, 1))
    for j in range(1, len(A)):
        E = np.einsum("ab,asc,bsd->cd", E, A[j], B[j])
    return float(E[0, 0])


def initial_w_sum(feat: np.ndarray, picks, jc: int, do_rel_cutoff=False):
    """fixedL.cc:702-728 given the drawn image positions `picks[label] = [n, ...]`.
    Returns W in the site-tensor layout of the rest of the oracle
    ([ml,d,mr], label site [ml,d,mr,NL])."""
    d = feat.shape[2]
    ipsis = []
    for lab in range(NL):
        psis = [mps_product_state(feat[n]) for n in picks[lab]]
        s = mps_sum(psis, 1e-10, 10, do_rel_cutoff)
        A = s[jc]
        T = np.zeros((A.shape[0], d, NL, A.shape[2]))
        T[:, :, lab, :] = 0.1 * A                       # Aref(c) *= 0.1*setElt(L(1+label))
        s[jc] = T.reshape(A.shape[0], d * NL, A.shape[2])
        ipsis.append(s)
    R = mps_sum(ipsis, 1e-8, 10, do_rel_cutoff)
    R[jc] = R[jc] / np.linalg.norm(R[jc])                # W.Aref(c) /= norm(W.A(c))
    W = [None] + [np.ascontiguousarray(R[j]) for j in range(1, len(R))]
    A = R[jc]
    W[jc] = np.ascontiguousarray(np.transpose(A.reshape(A.shape[0], d, NL, A.shape[2]), (0, 1, 3, 2)))
    return W
