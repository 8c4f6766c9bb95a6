"""Host-side data path of `fixedL`: MNIST idx reader with readMNIST's per-label
selection (mllib/mnist.h:443-530), the 2x2 block-mean `reduce` (image.h:316-346;
our `imglen` key), the feature map phi (fixedL.cc:637-642) and the
ITensor InputGroup file format (fixedL.cc:584-608)."""
from __future__ import annotations

import os
import re
import struct
import struct

import numpy as np

NL = 10


def _read_idx(path):
    with open(path, "rb") as f:
        raw = f.read()
    magic = struct.unpack(">I", raw[:4])[0]
    if magic == 2051:
        n, r, c = struct.unpack(">III", raw[4:16])
        return np.frombuffer(raw, np.uint8, n * r * c, 16).reshape(n, r * c)
    if magic == 2049:
        n = struct.unpack(">I", raw[4:8])[0]
        return np.frombuffer(raw, np.uint8, n, 8)
    raise ValueError(f"{path}: bad idx magic {magic}")


def readMNIST(datadir, kind="Train", NT=50000):
    """First NT images per label in file order; pixels / 255 (mnist.h:472-496)."""
    pre = "train" if kind == "Train" else "t10k"
    imgs = _read_idx(os.path.join(datadir, f"{pre}-images-idx3-ubyte"))
    labs = _read_idx(os.path.join(datadir, f"{pre}-labels-idx1-ubyte"))
    counts = [0] * NL
    keep = []
    for i, l in enumerate(labs.tolist()):
        if counts[l] >= NT:
            continue
        counts[l] += 1
        keep.append(i)
    keep = np.asarray(keep, np.int64)
    return imgs[keep].astype(np.float64) / 255.0, labs[keep].astype(np.int32)


def reduce(data, newlen):
    """image.h:316-346: block mean over bsize x bsize pixels, kept real."""
    n, npix = data.shape
    L = int(round(npix ** 0.5))
    if newlen == L:
        return data
    bs = L // newlen
    rem = L % bs
    img = data.reshape(n, L, L)[:, rem:rem + bs * newlen, rem:rem + bs * newlen]
    return img.reshape(n, newlen, bs, newlen, bs).mean(axis=(2, 4)).reshape(n, newlen * newlen)


def phi(g, d=2):
    """fixedL.cc:637-642: x = g/255 ; phi_n = (x/4)^(n-1)."""
    g = np.asarray(g, np.float64)
    if (g < 0).any() or (g > 255.0).any():
        raise ValueError("Expected g to be in [0,255]")
    x = g / 255.0
    return np.stack([(x / 4.0) ** n for n in range(d)], axis=-1)


class InputGroup:
    """`input { key = value ... }` (ITensor InputGroup): unknown keys are
    ignored, getYesNo accepts yes/no."""

    def __init__(self, path, name="input"):
        txt = open(path).read()
        m = re.search(r"\b" + re.escape(name) + r"\s*\{(.*?)\}", txt, re.S)
        if not m:
            raise ValueError(f"no group '{name}' in {path}")
        self.kv = {}
        for line in m.group(1).splitlines():
            line = line.split("#")[0].strip()
            if "=" in line:
                k, v = line.split("=", 1)
                self.kv[k.strip()] = v.strip()

    def getString(self, k, default=None):
        if k in self.kv:
            return self.kv[k]
        if default is None:
            raise KeyError(k)
        return default

    def getInt(self, k, default=None):
        return int(float(self.getString(k, None if default is None else str(default))))

    def getReal(self, k, default=None):
        return float(self.getString(k, None if default is None else repr(default)))

    def getYesNo(self, k, default=False):
        v = self.getString(k, "yes" if default else "no").lower()
        return v in ("yes", "y", "true", "1")


def synthetic_digits(NT, L=14, seed=20260925, first=0, nproto=10):
    """MNIST-shaped synthetic images for the GPU box (no dataset travels):
    10 smooth random 'stroke' prototypes on an LxL grid, per-image jitter and
    sparse noise, u8 pixels with MNIST-like sparsity (~20% non-zero);
    label = prototype id = image index % 10.  Counter-seeded per image so any
    shard [first, first+NT) of a larger set is reproducible.
    Returns (data[NT, L*L] float64 in [0,1] i.e. already /255, labels int32)."""
    yy, xx = np.mgrid[0:L, 0:L].astype(np.float64)
    prng = np.random.default_rng(seed)
    protos = np.zeros((nproto, L, L))
    for k in range(nproto):
        pts = prng.uniform(0.2 * L, 0.8 * L, size=(4, 2))
        for t in np.linspace(0, 1, 24):
            # piecewise-linear stroke through 4 control points
            seg = min(int(t * 3), 2)
            u = t * 3 - seg
            cx, cy = (1 - u) * pts[seg] + u * pts[seg + 1]
            protos[k] += np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * (0.055 * L) ** 2))
        protos[k] /= protos[k].max()
    out = np.empty((NT, L * L), np.float64)
    labels = np.empty(NT, np.int32)
    for i in range(NT):
        n = first + i
        r = np.random.default_rng([seed, n])
        k = n % nproto
        dx, dy = r.integers(-1, 2, size=2)
        img = np.roll(np.roll(protos[k], dx, axis=1), dy, axis=0) * r.uniform(0.6, 1.0)
        img = img + 0.08 * r.standard_normal((L, L)) * (r.random((L, L)) < 0.15)
        img = np.where(img > 0.25, img, 0.0)
        out[i] = np.clip(np.floor(img * 255.0), 0, 255).reshape(-1) / 255.0
        labels[i] = k
    return out, labels


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    return x ^ (x >> np.uint64(31))


def synthetic_pixels(NT, npix, seed=20260925, first=0):
    """BASELINE config 5 recipe: counter-based hash RNG splitmix64(seed, image, pixel) ->
    zero with probability 0.81, else uniform{1..255} (MNIST sparsity: 19.1 % non-zero); label = image
    mod 10.  Vectorised, any shard [first, first+NT) of the global set is reproducible.
    Returns (pixels / 255 as float64 [NT, npix], labels int32)."""
    with np.errstate(over="ignore"):
        img = (np.arange(NT, dtype=np.uint64) + np.uint64(first))[:, None]
        pixi = np.arange(npix, dtype=np.uint64)[None, :]
        h = _splitmix64(_splitmix64(np.uint64(seed) + img * np.uint64(1000003)) + pixi)
        u = (h >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
        v = _splitmix64(h)
        val = (v % np.uint64(255)).astype(np.float64) + 1.0
    out = np.where(u < 0.81, 0.0, val) / 255.0
    labels = ((np.arange(NT, dtype=np.int64) + first) % NL).astype(np.int32)
    return out, labels


def window_mps(N, d=2, m=300, seed=5, jc=None):
    """Short chain with the SAME link dimension m on every interior bond (config-5 window, SURVEY 8d:
    the kernels at m_l = m_r = maxm are measured on a few sites instead of a 196-site chain whose
    environments would not fit): i.i.d. N(0, 1/(m_a d)) site tensors with a near-identity s=0 slice,
    label index on site N//2, centre tensors normalised."""
    jc = N // 2 if jc is None else jc
    rng = np.random.default_rng(seed)
    dims = [1] + [m] * (N - 1) + [1]
    W = [None] * (N + 1)
    for j in range(1, N + 1):
        ml, mr = dims[j - 1], dims[j]
        shape = (ml, d, mr, NL) if j == jc else (ml, d, mr)
        A = rng.standard_normal(shape) / np.sqrt(ml * d)
        if j == jc:
            A[:, 0, :, :] += np.eye(ml, mr)[:, :, None]
        else:
            A[:, 0, :] += np.eye(ml, mr)
        W[j] = A / np.linalg.norm(A) * np.sqrt(min(ml, mr))
    W[jc] = W[jc] / np.linalg.norm(W[jc])
    return W


def random_mps(N, d=2, m=10, seed=1, noise=0.3, jc=None):
    """Deterministic start MPS (replaces the time-seeded init of
    fixedL.cc:702-728): near-identity s=0 slices so environments stay
    O(1), right-canonical with the centre on site 1, label index on site N//2,
    W[jc] /= norm (fixedL.cc:725).  Returns a list W[0..N] (W[0] unused)."""
    jc = N // 2 if jc is None else jc
    rng = np.random.default_rng(seed)
    dims = [1] * (N + 1)
    for j in range(1, N):
        cap = min(j, N - j)
        dims[j] = m if cap >= 40 else min(m, d ** cap)
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
    for j in range(N, 1, -1):
        A = W[j]
        ml = A.shape[0]
        Q, R = np.linalg.qr(A.reshape(ml, -1).T)
        k = Q.shape[1]
        W[j] = np.ascontiguousarray(Q.T.reshape((k,) + A.shape[1:]))
        R = R / np.linalg.norm(R) * np.sqrt(k)
        W[j - 1] = np.ascontiguousarray(np.moveaxis(np.tensordot(W[j - 1], R.T, axes=([2], [0])), -1, 2))
    W[1] = W[1] / np.linalg.norm(W[1])
    W[jc] = W[jc] / np.linalg.norm(W[jc])
    return W


# ---- files shared with the C++ host program (tnml_b200/host/itensor_lite.h) -----
def _w_index(f, idx_id, m, name, typ):
    nb, tb = name.encode(), typ.encode()
    f.write(struct.pack("<qqq", idx_id, m, len(nb)) + nb + struct.pack("<q", len(tb)) + tb)


def write_sites_file(path, N, d=2):
    """`sites` file in the TNMLS1 format read by fixedL (SiteSet::read)."""
    with open(path, "wb") as f:
        f.write(b"TNMLS1\0\0" + struct.pack("<q", N))
        for j in range(1, N + 1):
            _w_index(f, j, d, f"S{j}", "Site")


def write_mps_file(path, W, d=2):
    """`W` file in the TNMLW1 format (index order left, site, right [, L])."""
    N = len(W) - 1
    with open(path, "wb") as f:
        f.write(b"TNMLW1\0\0" + struct.pack("<q", N))
        for j in range(1, N + 1):
            A = np.ascontiguousarray(W[j], np.float64)
            f.write(struct.pack("<q", A.ndim))
            _w_index(f, 1000 + j - 1, A.shape[0], f"l{j - 1}", "Link")
            _w_index(f, j, d, f"S{j}", "Site")
            _w_index(f, 1000 + j, A.shape[2], f"l{j}", "Link")
            if A.ndim == 4:
                _w_index(f, 5000, NL, "L", "Label")
            f.write(struct.pack("<q", A.size) + A.tobytes())


def _r_index(f):
    idx_id, m, n = struct.unpack("<qqq", f.read(24))
    name = f.read(n).decode()
    (tn,) = struct.unpack("<q", f.read(8))
    typ = f.read(tn).decode()
    return idx_id, m, name, typ


def read_mps_file(path):
    """Read a `W` file written by the fixedL program (TNMLW1): returns the 1-indexed list of site
    tensors [ml,d,mr] (label site [ml,d,mr,NL])."""
    with open(path, "rb") as f:
        if f.read(8)[:6] != b"TNMLW1":
            raise ValueError(f"not a tnml_b200 MPS file: {path}")
        (N,) = struct.unpack("<q", f.read(8))
        W = [None]
        for _ in range(N):
            (r,) = struct.unpack("<q", f.read(8))
            dims = [_r_index(f)[1] for _ in range(r)]
            (n,) = struct.unpack("<q", f.read(8))
            W.append(np.frombuffer(f.read(8 * n), np.float64).reshape(dims).copy())
    return W


def write_idx_files(datadir, pix_u8, labels, side, kind="train"):
    """MNIST-format idx3/idx1 files (for running the fixedL binary on synthetic data)."""
    os.makedirs(datadir, exist_ok=True)
    n = pix_u8.shape[0]
    with open(os.path.join(datadir, f"{kind}-images-idx3-ubyte"), "wb") as f:
        f.write(struct.pack(">IIII", 2051, n, side, side) + np.ascontiguousarray(pix_u8, np.uint8).tobytes())
    with open(os.path.join(datadir, f"{kind}-labels-idx1-ubyte"), "wb") as f:
        f.write(struct.pack(">II", 2049, n) + np.ascontiguousarray(labels, np.uint8).tobytes())
