"""Developer study (CPU, scipy): PCG iteration counts on the reduced camera system for candidate preconditioners.

Iteration counts are arithmetic, not hardware: they can be measured here and carried to the B200 kernel (csrc/pcg.cuh), where one
iteration costs ~113 us on cfg5 (70 us of it the S.p product at the HBM roofline).  The reduced system S = Hpp + lambda I - W Hll^-1 W^T
is assembled with scipy from the CPU checker's blocks (oracle.pyoracle.ba_build — used as a source of test matrices, nothing here is product).

  python tools/pcg_precond_study.py [K] [P] [obs_per_point] [window]

Preconditioners compared (all SPD, all applied inside the same scipy CG loop, tolerance ||r|| <= 1e-8 ||b||):
  bj        block-Jacobi: inverse of the 6x6 diagonal blocks                         (round-0 kernel)
  bj+c      bj + piecewise-constant coarse correction over NC aggregates             (the kernel's current two-level preconditioner)
  seg(m)    non-overlapping additive Schwarz: exact inverse of each diagonal block of m consecutive keyframes
  seg(m)+c  seg(m) + the same coarse correction
  ovl(m,o)  restricted-overlap variant: blocks of m keyframes extended by o on each side, contributions summed (symmetric additive Schwarz)
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccm_slam_b200 import synth  # noqa: E402
from oracle import pyoracle  # noqa: E402


def reduced_system(p, lam):
    b = pyoracle.ba_build(p)
    K, P, E = p.K, p.P, p.E
    free = np.flatnonzero(np.asarray(p.fixed) == 0)
    slot = -np.ones(K, np.int64); slot[free] = np.arange(len(free))
    Kf = len(free)
    kf = np.asarray(p.obs_kf); mp = np.asarray(p.obs_mp)
    keep = slot[kf] >= 0
    # W as BSR (Kf x P blocks of 6x3)
    order = np.lexsort((mp[keep], slot[kf[keep]]))
    rows = slot[kf[keep]][order]; cols = mp[keep][order]
    indptr = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=Kf))])
    W = sp.bsr_matrix((b["W"][keep][order], cols, indptr), shape=(6 * Kf, 3 * P)).tocsr()
    Hll = b["Hll"] + lam * np.eye(3)[None]
    Dinv = sp.bsr_matrix((np.linalg.inv(Hll), np.arange(P), np.arange(P + 1)), shape=(3 * P, 3 * P)).tocsr()
    Hpp = sp.bsr_matrix((b["Hpp"][free] + lam * np.eye(6)[None], np.arange(Kf), np.arange(Kf + 1)), shape=(6 * Kf, 6 * Kf)).tocsr()
    WD = W @ Dinv
    S = (Hpp - WD @ W.T).tocsr()
    rhs = b["bp"][free].reshape(-1) - WD @ b["bl"].reshape(-1)
    return S, rhs, Kf, float(max(b["Hpp"][free].reshape(len(free), 36)[:, ::7].max(), b["Hll"].reshape(P, 9)[:, ::4].max()))


def diag_blocks(S, Kf, m):
    """dense inverses of the diagonal blocks of m keyframes (last one may be shorter)"""
    out = []
    for s in range(0, Kf, m):
        e = min(Kf, s + m)
        A = S[6 * s:6 * e, 6 * s:6 * e].toarray()
        out.append((6 * s, 6 * e, np.linalg.inv(A)))
    return out


def make_seg(S, Kf, m, overlap=0):
    if overlap == 0:
        blocks = diag_blocks(S, Kf, m)
    else:
        blocks = []
        for s in range(0, Kf, m):
            a, e = max(0, s - overlap), min(Kf, s + m + overlap)
            blocks.append((6 * a, 6 * e, np.linalg.inv(S[6 * a:6 * e, 6 * a:6 * e].toarray())))

    def apply(r):
        z = np.zeros_like(r)
        for a, e, Ainv in blocks:
            z[a:e] += Ainv @ r[a:e]
        return z
    nbytes = sum(Ainv.nbytes for _, _, Ainv in blocks)
    return apply, nbytes


def prolongation(Kf, nc_target, kind):
    """6 coarse unknowns per aggregate of consecutive keyframes.  kind: "pc" piecewise constant (the kernel's), "pl" piecewise
    linear between aggregate centres (hat functions), applied per degree of freedom."""
    agg = -(-Kf // nc_target)            # keyframes per aggregate
    nagg = -(-Kf // agg)
    k = np.arange(Kf)
    if kind == "pc":
        r = np.repeat(k, 1); c = k // agg; w = np.ones(Kf)
    else:
        centre = (np.arange(nagg) + 0.5) * agg - 0.5
        pos = np.clip((k - centre[0]) / agg, 0, nagg - 1)
        lo = np.minimum(np.floor(pos).astype(np.int64), nagg - 2) if nagg > 1 else np.zeros(Kf, np.int64)
        f = np.clip(pos - lo, 0.0, 1.0)
        r = np.concatenate([k, k]); c = np.concatenate([lo, np.minimum(lo + 1, nagg - 1)]); w = np.concatenate([1 - f, f])
    Pk = sp.csr_matrix((w, (r, c)), shape=(Kf, nagg))
    return sp.kron(Pk, sp.identity(6), format="csr"), 6 * nagg


def make_coarse(S, Kf, nc_target, kind="pc", smooth=0.0, Dinv=None):
    Pm, nc = prolongation(Kf, nc_target, kind)
    if smooth > 0.0:                      # smoothed aggregation: one damped block-Jacobi sweep on the prolongator
        Pm = (Pm - smooth * (Dinv @ (S @ Pm))).tocsr()
    Ac = (Pm.T @ S @ Pm).toarray()
    Acinv = np.linalg.inv(Ac)
    return (lambda r: Pm @ (Acinv @ (Pm.T @ r))), nc, Pm.nnz


def pcg(S, b, M, tol=1e-8, maxit=5000):
    x = np.zeros_like(b); r = b.copy(); z = M(r); p = z.copy()
    rz = r @ z; bn = np.linalg.norm(b)
    for it in range(1, maxit + 1):
        q = S @ p
        a = rz / (p @ q)
        x += a * p; r -= a * q
        if np.linalg.norm(r) <= tol * bn:
            return it
        z = M(r); rz2 = r @ z
        p = z + (rz2 / rz) * p; rz = rz2
    return maxit


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
    opp = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    win = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    t = time.time()
    p = synth.make_global_ba(K=K, P=P, obs_per_point=opp, window=win, seed=4, name="study")
    print(f"problem K={p.K} P={p.P} E={p.E} ({time.time() - t:.1f}s)", flush=True)
    S0, rhs, Kf, maxdiag = reduced_system(p, 0.0)
    lam0 = 1e-5 * maxdiag
    print(f"S: {S0.shape[0]} unknowns, {S0.nnz / 36:.0f} blocks; lambda0 = {lam0:.3e}", flush=True)
    lams = [lam0 / 9, lam0 / 81, lam0 / 729, lam0 / 6561]
    for lam in lams:
        S, rhs, Kf, _ = reduced_system(p, lam)
        res = {}
        bj, _ = make_seg(S, Kf, 1)
        blocks = diag_blocks(S, Kf, 1)
        Dinv = sp.block_diag([B for _, _, B in blocks], format="csr")
        res["bj"] = pcg(S, rhs, bj)
        for NC in (192, 384, 768):
            for kind in ("pc", "pl"):
                coarse, nc, _ = make_coarse(S, Kf, NC, kind)
                res[f"bj+{kind}{NC}"] = pcg(S, rhs, lambda r: bj(r) + coarse(r))
            coarse, nc, nnz = make_coarse(S, Kf, NC, "pc", smooth=0.66, Dinv=Dinv)
            res[f"bj+sa{NC}"] = pcg(S, rhs, lambda r: bj(r) + coarse(r))
            res[f"sa{NC} P nnz/row"] = round(nnz / (6 * Kf), 1)
        print(f"lambda = {lam:.3e}: " + "  ".join(f"{k}={v}" for k, v in res.items()), flush=True)


if __name__ == "__main__":
    main()
