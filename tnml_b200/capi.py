"""ctypes binding of the C-ABI in include/tnml_b200.h.

Only loads the in-tree shared library `tnml_b200/libtnml_b200.so` (built by
`__graft_entry__.build()` / `make`).  There is no Python or CPU fallback: if
the library is missing, or no CUDA device is present when a handle is
created, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

NL = 10
D = 2
FROMLEFT = 1
FROMRIGHT = 2
UNIQUE_ID_BYTES = 128

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtnml_b200.so")


class TnmlError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tnml_b200 error {code}: {msg}")
        self.code = code


class BondParams(C.Structure):
    _fields_ = [("Npass", C.c_int), ("lambda_", C.c_double), ("cconv", C.c_double),
                ("cutoff", C.c_double), ("maxm", C.c_int), ("minm", C.c_int),
                ("do_rel_cutoff", C.c_int)]


class BondResult(C.Structure):
    _fields_ = [("origm", C.c_int), ("newm", C.c_int), ("truncerr", C.c_double),
                ("cost", C.c_double), ("cost_label", C.c_double * NL),
                ("ncorrect", C.c_int64), ("normB", C.c_double), ("dB", C.c_double),
                ("npass_done", C.c_int), ("cg_cost", C.c_double * 8),
                ("cg_rnorm", C.c_double * 8), ("svd_sweeps", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("launches", C.c_int64), ("alg_bytes", C.c_double), ("alg_flops", C.c_double),
                ("ms_proj", C.c_double), ("ms_grad", C.c_double), ("ms_fat", C.c_double),
                ("ms_svd", C.c_double), ("ms_shift", C.c_double), ("ms_other", C.c_double),
                ("tier_evictions", C.c_int64), ("tier_fetches", C.c_int64), ("tier_bytes", C.c_double)]


# every symbol include/tnml_b200.h declares (tests check the library exports all)
SYMBOLS = [
    "tnml_version", "tnml_last_error", "tnml_create", "tnml_destroy", "tnml_set_images",
    "tnml_set_site", "tnml_get_site_dims", "tnml_get_site", "tnml_init_envs", "tnml_set_bond",
    "tnml_bond_form", "tnml_bond_dims", "tnml_bond_load", "tnml_bond_store", "tnml_cgrad",
    "tnml_svd_split", "tnml_quadcost", "tnml_shift_env", "tnml_bond_update", "tnml_predict", "tnml_fulltest",
    "tnml_get_env", "tnml_comm_get_unique_id", "tnml_comm_init_rank", "tnml_comm_broadcast", "tnml_set_option", "tnml_get_stats",
    "tnml_set_timing", "tnml_synchronize", "tnml_stream",
]

_lib = None


def load_library():
    """Load (once) and type the C-ABI.  Raises if the extension is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not built -- run `python -c 'import __graft_entry__ as g; g.build()'`"
                          " (no CPU fallback exists)")
    lib = C.CDLL(LIB_PATH)
    vp, i, d, i64 = C.c_void_p, C.c_int, C.c_double, C.c_int64
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.tnml_version.restype = C.c_char_p
    lib.tnml_last_error.restype = C.c_char_p
    lib.tnml_last_error.argtypes = [vp]
    lib.tnml_create.argtypes = [i, i, C.POINTER(vp)]
    lib.tnml_destroy.argtypes = [vp]
    lib.tnml_set_images.argtypes = [vp, i64, i, vp, vp, i64, i64]
    lib.tnml_set_site.argtypes = [vp, i, i, i, i, vp]
    lib.tnml_get_site_dims.argtypes = [vp, i, ip, ip, ip]
    lib.tnml_get_site.argtypes = [vp, i, vp, C.c_size_t]
    lib.tnml_init_envs.argtypes = [vp]
    lib.tnml_set_bond.argtypes = [vp, i]
    lib.tnml_bond_form.argtypes = [vp]
    lib.tnml_bond_dims.argtypes = [vp, ip, ip, ip]
    lib.tnml_bond_load.argtypes = [vp, vp, C.c_size_t]
    lib.tnml_bond_store.argtypes = [vp, vp, C.c_size_t]
    lib.tnml_cgrad.argtypes = [vp, i, d, d, dp, dp, ip]
    lib.tnml_svd_split.argtypes = [vp, i, d, i, i, i, ip, dp]
    lib.tnml_quadcost.argtypes = [vp, i, d, dp, dp, C.POINTER(i64)]
    lib.tnml_shift_env.argtypes = [vp, i, i]
    lib.tnml_bond_update.argtypes = [vp, i, i, C.POINTER(BondParams), C.POINTER(BondResult)]
    lib.tnml_predict.argtypes = [vp, vp, vp]
    lib.tnml_fulltest.argtypes = [vp, vp, C.POINTER(i64)]
    lib.tnml_get_env.argtypes = [vp, i, ip, ip, vp, C.c_size_t]
    lib.tnml_comm_get_unique_id.argtypes = [vp]
    lib.tnml_comm_init_rank.argtypes = [vp, i, i, vp]
    lib.tnml_comm_broadcast.argtypes = [vp, dp, i, i]
    lib.tnml_set_option.argtypes = [vp, C.c_char_p, d]
    lib.tnml_get_stats.argtypes = [vp, C.POINTER(Stats), i]
    lib.tnml_set_timing.argtypes = [vp, i]
    lib.tnml_synchronize.argtypes = [vp]
    lib.tnml_stream.argtypes = [vp]
    lib.tnml_stream.restype = vp
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("tnml_version", "tnml_last_error", "tnml_stream"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Handle:
    """Owns one `tnml_handle` (one GPU)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self._h = C.c_void_p()
        rc = self.lib.tnml_create(device, 0, C.byref(self._h))
        if rc != 0:
            raise TnmlError(rc, self.lib.tnml_last_error(None).decode())

    def close(self):
        if self._h:
            self.lib.tnml_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise TnmlError(rc, self.lib.tnml_last_error(self._h).decode())

    # ---- data -----------------------------------------------------------
    def set_images(self, feat: np.ndarray, labels: np.ndarray, NT_global: int = 0, first: int = 0):
        feat = np.ascontiguousarray(feat, np.float64)
        labels = np.ascontiguousarray(labels, np.int32)
        NT, N, d = feat.shape
        assert d == D and labels.shape == (NT,)
        self.NT, self.N, self.jc = NT, N, N // 2
        self._ck(self.lib.tnml_set_images(self._h, NT, N, _ptr(feat), _ptr(labels), NT_global or NT, first))

    def set_site(self, j: int, A: np.ndarray):
        A = np.ascontiguousarray(A, np.float64)
        has_label = int(A.ndim == 4)
        self._ck(self.lib.tnml_set_site(self._h, j, A.shape[0], A.shape[2], has_label, _ptr(A)))

    def get_site(self, j: int) -> np.ndarray:
        ml, mr, lab = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.lib.tnml_get_site_dims(self._h, j, C.byref(ml), C.byref(mr), C.byref(lab)))
        shape = (ml.value, D, mr.value, NL) if lab.value else (ml.value, D, mr.value)
        out = np.empty(shape, np.float64)
        self._ck(self.lib.tnml_get_site(self._h, j, _ptr(out), out.size))
        return out

    def set_mps(self, W):
        for j in range(1, self.N + 1):
            self.set_site(j, W[j])

    def get_mps(self):
        return [None] + [self.get_site(j) for j in range(1, self.N + 1)]

    # ---- phases ---------------------------------------------------------
    def init_envs(self):
        self._ck(self.lib.tnml_init_envs(self._h))

    def set_bond(self, b: int):
        self._ck(self.lib.tnml_set_bond(self._h, b))

    def bond_form(self):
        self._ck(self.lib.tnml_bond_form(self._h))

    def bond_shape(self):
        ml, mr, lab = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.lib.tnml_bond_dims(self._h, C.byref(ml), C.byref(mr), C.byref(lab)))
        return (ml.value, D, D, mr.value, NL) if lab.value else (ml.value, D, D, mr.value)

    def bond_load(self, B: np.ndarray):
        B = np.ascontiguousarray(B, np.float64)
        self._ck(self.lib.tnml_bond_load(self._h, _ptr(B), B.size))

    def bond_store(self) -> np.ndarray:
        out = np.empty(self.bond_shape(), np.float64)
        self._ck(self.lib.tnml_bond_store(self._h, _ptr(out), out.size))
        return out

    def cgrad(self, Npass=4, lam=0.0, cconv=1e-10):
        costs = (C.c_double * 8)()
        rn = (C.c_double * 8)()
        nd = C.c_int()
        self._ck(self.lib.tnml_cgrad(self._h, Npass, lam, cconv, costs, rn, C.byref(nd)))
        k = min(nd.value, 8)
        return list(costs[:k]), list(rn[:k])

    def svd_split(self, direction, cutoff, maxm, minm, do_rel_cutoff=False):
        m, te = C.c_int(), C.c_double()
        self._ck(self.lib.tnml_svd_split(self._h, direction, cutoff, maxm, minm, int(do_rel_cutoff),
                                         C.byref(m), C.byref(te)))
        return m.value, te.value

    def quadcost(self, use_sites=False, lam=0.0):
        c = C.c_double()
        cl = (C.c_double * NL)()
        nc = C.c_int64()
        self._ck(self.lib.tnml_quadcost(self._h, int(use_sites), lam, C.byref(c), cl, C.byref(nc)))
        return c.value, np.array(cl[:]), nc.value

    def shift_env(self, b, direction):
        self._ck(self.lib.tnml_shift_env(self._h, b, direction))

    def bond_update(self, b, ha, params: BondParams) -> BondResult:
        res = BondResult()
        self._ck(self.lib.tnml_bond_update(self._h, b, ha, C.byref(params), C.byref(res)))
        return res

    def predict(self, want_P=False):
        lab = np.empty(self.NT, np.int32)
        P = np.empty((self.NT, NL), np.float64) if want_P else None
        self._ck(self.lib.tnml_predict(self._h, _ptr(lab), _ptr(P) if want_P else None))
        return (lab, P) if want_P else lab

    def fulltest(self):
        """fullTest (util.h:123-200): returns (predicted labels, ncorrect)."""
        pred = np.empty(self.NT, np.int32)
        nc = C.c_int64()
        self._ck(self.lib.tnml_fulltest(self._h, _ptr(pred), C.byref(nc)))
        return pred, nc.value

    def get_env(self, slot):
        m, fat = C.c_int(), C.c_int()
        self._ck(self.lib.tnml_get_env(self._h, slot, C.byref(m), C.byref(fat), None, 0))
        shape = (self.NT, NL, m.value) if fat.value else (self.NT, m.value)
        out = np.empty(shape, np.float64)
        self._ck(self.lib.tnml_get_env(self._h, slot, C.byref(m), C.byref(fat), _ptr(out), out.size))
        return out

    # ---- comm / stats ---------------------------------------------------
    def comm_init_rank(self, nranks, rank, uid: bytes):
        buf = (C.c_uint8 * UNIQUE_ID_BYTES).from_buffer_copy(uid)
        self._ck(self.lib.tnml_comm_init_rank(self._h, nranks, rank, buf))

    def comm_broadcast(self, vals, root=0):
        buf = (C.c_double * len(vals))(*[float(v) for v in vals])
        self._ck(self.lib.tnml_comm_broadcast(self._h, buf, len(vals), root))
        return list(buf)

    def set_option(self, name: str, value: float):
        self._ck(self.lib.tnml_set_option(self._h, name.encode(), float(value)))

    def stats(self, reset=False) -> Stats:
        s = Stats()
        self._ck(self.lib.tnml_get_stats(self._h, C.byref(s), int(reset)))
        return s

    def set_timing(self, on=True):
        self._ck(self.lib.tnml_set_timing(self._h, int(on)))

    def synchronize(self):
        self._ck(self.lib.tnml_synchronize(self._h))

    def stream(self) -> int:
        return int(self.lib.tnml_stream(self._h) or 0)


def comm_get_unique_id() -> bytes:
    lib = load_library()
    buf = (C.c_uint8 * UNIQUE_ID_BYTES)()
    rc = lib.tnml_comm_get_unique_id(buf)
    if rc != 0:
        raise TnmlError(rc, lib.tnml_last_error(None).decode())
    return bytes(buf)
