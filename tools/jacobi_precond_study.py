"""CPU study behind the SVD preconditioner choice: number of one-sided Jacobi
sweeps on R^T after one QR ("qr") vs column sort + two QRs ("sqr2"), for the cyclic ordering and
for the block orderings the device kernel uses (blk8full == jacobi_gram_kernel<8>).  The block
simulation reproduces the sweep counts measured on the GPU (tools/svd_bench.py).
  python tools/jacobi_precond_study.py [n]"""
import sys, numpy as np
EPS = 1.1102230246251565e-16
def rr_pair(n, r, k):
    if k == 0: return n-1, r
    return (r+k) % (n-1), (r-k) % (n-1)
def rotate(A, Jm, p, q, tol):
    xp, xq = A[:, p], A[:, q]
    a = (xp*xp).sum(0); b = (xq*xq).sum(0); g = (xp*xq).sum(0)
    ab = a*b
    doit = (ab > 0) & (g*g > tol*tol*ab)
    if not doit.any(): return 0.0
    mx = np.sqrt((g*g/np.where(ab>0,ab,1))[doit].max())
    d = b-a; h = 2*g
    r2 = np.where(doit, d*d+h*h, 1.0); rinv = 1/np.sqrt(r2)
    c2 = 0.5*np.abs(d)*rinv+0.5; c = np.sqrt(c2); s = np.copysign(0.5,d)*h*rinv/c
    c = np.where(doit,c,1.0); s = np.where(doit,s,0.0)
    A[:, p] = c*xp - s*xq; A[:, q] = s*xp + c*xq
    return mx
def sweep_block(A, W, tol, mode="full"):
    n = A.shape[1]; nblk = (n+W-1)//W; nblk_e = nblk + (nblk % 2)
    assert n % W == 0
    mx = 0.0
    for R in range(nblk_e-1):
        prs = [rr_pair(nblk_e, R, k) for k in range(nblk_e//2)]
        prs = [(P,Q) for (P,Q) in prs if P < nblk and Q < nblk]
        cols = np.array([list(range(P*W,(P+1)*W)) + list(range(Q*W,(Q+1)*W)) for (P,Q) in prs])
        if mode == "full" or R == 0:
            for rd in range(2*W-1):
                lp = [rr_pair(2*W, rd, k) for k in range(W)]
                p = cols[:, [x[0] for x in lp]].ravel(); q = cols[:, [x[1] for x in lp]].ravel()
                mx = max(mx, rotate(A, None, p, q, tol))
        else:
            for rd in range(W):
                p = cols[:, list(range(W))].ravel(); q = cols[:, [W + (k+rd) % W for k in range(W)]].ravel()
                mx = max(mx, rotate(A, None, p, q, tol))
    return mx
def sweep_cyclic(A, tol):
    n = A.shape[1]; mx = 0.0
    for r in range(n-1):
        lp = [rr_pair(n, r, k) for k in range(n//2)]
        mx = max(mx, rotate(A, None, np.array([x[0] for x in lp]), np.array([x[1] for x in lp]), tol))
    return mx
def count(A, fn, stop=1e-8):
    A = A.copy(); tol = 8*np.sqrt(A.shape[0])*EPS
    h = []
    for sw in range(40):
        mx = fn(A, tol); h.append(mx)
        if mx <= stop: break
    return h
def precond(M, kind):
    if kind == "qr": return np.linalg.qr(M)[1].T
    nrm = np.linalg.norm(M, axis=0); Ms = M[:, np.argsort(-nrm)]
    R = np.linalg.qr(Ms)[1]
    if kind == "sqr": return R.T
    R2 = np.linalg.qr(R.T)[1]
    if kind == "sqr2": return R2.T
    R3 = np.linalg.qr(R2.T)[1]
    if kind == "sqr3": return R3.T
def cases(n=240, seed=0):
    rng = np.random.default_rng(seed)
    m = n//2
    A1 = rng.standard_normal((n, m))/np.sqrt(n); A2 = rng.standard_normal((m, n))/np.sqrt(m)
    B0 = A1 @ A2
    out = {}
    out["lowrank+1e-3"] = B0 + 1e-3*np.linalg.norm(B0)/n*rng.standard_normal((n,n))
    out["random"] = rng.standard_normal((n,n))
    U = rng.standard_normal((n,n)); V = rng.standard_normal((n,n))
    out["graded1e-9"] = (U*np.logspace(0,-9,n)) @ V
    return out
if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 240
    for name, M in cases(n).items():
        for kind in ("qr", "sqr2"):
            A = precond(M, kind)
            for label, fn in (("cyclic", sweep_cyclic), ("blk8full", lambda A,t: sweep_block(A,8,t,"full")),
                              ("blk8cross", lambda A,t: sweep_block(A,8,t,"cross")), ("blk16cross", lambda A,t: sweep_block(A,16,t,"cross"))):
                h = count(A, fn)
                print(f"{name:14s} {kind:5s} {label:10s} sweeps {len(h)}  " + " ".join(f"{x:.0e}" for x in h), flush=True)
