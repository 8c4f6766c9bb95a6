"""Driver of oracle/fixedl_ref_cpu.cpp (the threaded C++ literal restatement).
TEST / BASELINE INFRASTRUCTURE ONLY.  Builds into oracle/_build/."""
import os
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "_build", "fixedl_ref_cpu")


def build():
    src = os.path.join(HERE, "fixedl_ref_cpu.cpp")
    if os.path.exists(BIN) and os.path.getmtime(BIN) >= os.path.getmtime(src):
        return BIN
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    subprocess.run(["g++", "-O3", "-march=native", "-std=c++17", "-pthread", "-o", BIN, src], check=True)
    return BIN


def write_problem(path, x, y, l, r, labels, B, Npass=4, lam=0.0, cconv=1e-10):
    """l: [NT,ml] or [NT,NL,ml] (fat); r likewise; B [ml,2,2,mr(,NL)]."""
    NT = x.shape[0]
    cls = 1 if B.ndim == 5 else (0 if r.ndim == 3 else 2)
    ml, mr = B.shape[0], B.shape[3]
    with open(path, "wb") as f:
        f.write(struct.pack("<5q2d", NT, ml, mr, cls, Npass, lam, cconv))
        for a in (x, y, l, r):
            f.write(np.ascontiguousarray(a, np.float64).tobytes())
        f.write(np.ascontiguousarray(labels, np.int32).tobytes())
        f.write(np.ascontiguousarray(B, np.float64).tobytes())


def run(problem, result, nthread=1, reps=1, Bshape=None):
    build()
    out = subprocess.run([BIN, problem, result, str(nthread), str(reps)], check=True, capture_output=True, text=True)
    raw = open(result, "rb").read()
    nc = struct.unpack_from("<q", raw, 0)[0]
    costs = np.frombuffer(raw, np.float64, nc, 8)
    off = 8 + 8 * nc
    C = struct.unpack_from("<d", raw, off)[0]
    ncor = struct.unpack_from("<q", raw, off + 8)[0]
    tt = np.frombuffer(raw, np.float64, 3, off + 16)
    B = np.frombuffer(raw, np.float64, -1, off + 40)
    if Bshape is not None:
        B = B.reshape(Bshape)
    return dict(costs=costs, C=C, ncor=ncor, t_setbond=tt[0], t_cgrad=tt[1], t_quadcost=tt[2], B=B,
                log=out.stdout.strip())


def problem_from_oracle(ts, B):
    """Current bond of an oracle TrainStates -> arrays for write_problem."""
    b = ts.currb
    LE, RE = ts.envs(b)
    NT = ts.NT
    l = np.ones((NT, 1)) if LE is None else LE
    r = np.ones((NT, 1)) if RE is None else RE
    return ts.feat[:, b - 1], ts.feat[:, b], l, r, ts.labels, B
