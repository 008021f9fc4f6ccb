"""CPU study (numpy/scipy, no GPU): would ONE reduction per iteration cost PCG iterations or accuracy on the reduced camera systems?

The distributed solve (k_pcg2) exchanges scalars twice per iteration (p.q after the product, r.z / r.r after the
preconditioner).  Chronopoulos-Gear's rearrangement of preconditioned CG needs one: with u = M^-1 r and w = S u,
gamma = r.u and delta = w.u are reduced together, beta = gamma / gamma_old, alpha = gamma / (delta - beta gamma / alpha_old).
This script builds S and b of the first LM trial with the oracle (cfg5 at 1/10 trajectory length by default, lambda as the LM loop
starts it), the two-level preconditioner of k_pcg2 (block-Jacobi + piecewise-linear coarse space, csrc/pcg.cuh coarse_parents), and
runs both recurrences to the same stopping rule (||r|| <= tol ||b||), reporting iterations and the TRUE residual at exit.

  python tools/cg_variants_cpu.py [config] [K_div] [nc]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import scipy.sparse as sp  # noqa: E402
from ccm_slam_b200 import synth  # noqa: E402
from oracle import pyoracle  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
div = int(sys.argv[2]) if len(sys.argv) > 2 else 10
NC = int(sys.argv[3]) if len(sys.argv) > 3 else 128
cfg = synth.CONFIGS[name]
p = synth.make_config(name, K=cfg["K"] // div, P=cfg["P"] // div) if div > 1 else synth.make_config(name)
print(f"[{name}/{div}] K={p.K} P={p.P} E={p.E}", flush=True)

t0 = time.time()
blk = pyoracle.ba_build(p, huber_delta=float(np.float32(np.sqrt(5.99))))
free = np.flatnonzero(~p.fixed.astype(bool))
slot = -np.ones(p.K, int); slot[free] = np.arange(free.size)
Kf = free.size
# lambda as g2o's LM starts it: tau * max diagonal of the Hessian (tau = 1e-5)
lam = 1e-5 * max(np.abs(np.einsum("kii->ki", blk["Hpp"])[free]).max(), np.abs(np.einsum("kii->ki", blk["Hll"])).max())
lam0 = lam


def run(lam):
    t0 = time.time()
    Vinv = np.linalg.inv(blk["Hll"] + lam * np.eye(3))
    okf, omp = p.obs_kf, p.obs_mp
    m = slot[okf] >= 0
    rows = slot[okf[m]]; lms = omp[m]; W = blk["W"][m]                        # (E', 6, 3)
    # W as a sparse (6 Kf) x (3 P) matrix, S = Hpp + lam I - W Vinv W^T, b = bp - W Vinv bl
    E_ = rows.size
    ri = (rows[:, None, None] * 6 + np.arange(6)[None, :, None]).repeat(3, axis=2).ravel()
    ci = (lms[:, None, None] * 3 + np.arange(3)[None, None, :]).repeat(6, axis=1).ravel()
    Wm = sp.csr_matrix((W.ravel(), (ri, ci)), shape=(6 * Kf, 3 * p.P))
    pi = (np.arange(p.P)[:, None, None] * 3 + np.arange(3)[None, :, None]).repeat(3, axis=2).ravel()
    pj = (np.arange(p.P)[:, None, None] * 3 + np.arange(3)[None, None, :]).repeat(3, axis=1).ravel()
    Vb = sp.csr_matrix((Vinv.ravel(), (pi, pj)), shape=(3 * p.P, 3 * p.P))
    hi = (np.arange(Kf)[:, None, None] * 6 + np.arange(6)[None, :, None]).repeat(6, axis=2).ravel()
    hj = (np.arange(Kf)[:, None, None] * 6 + np.arange(6)[None, None, :]).repeat(6, axis=1).ravel()
    Hd = sp.csr_matrix(((blk["Hpp"][free] + lam * np.eye(6)).ravel(), (hi, hj)), shape=(6 * Kf, 6 * Kf))
    S = (Hd - Wm @ Vb @ Wm.T).tocsr()
    b = blk["bp"][free].ravel() - Wm @ (Vb @ blk["bl"].ravel())
    print(f"S: {S.shape[0]} unknowns, {S.nnz / 36:.0f} blocks, lambda {lam:.3e}, built in {time.time() - t0:.1f}s", flush=True)

    # two-level preconditioner: blockdiag(S)^-1 + P (P^T S P)^-1 P^T, piecewise-linear P over nc nodes (csrc/pcg.cuh coarse_parents)
    Sd = S.toarray() if S.shape[0] <= 8000 else None
    Dinv = np.stack([np.linalg.inv(S[6 * a:6 * a + 6, 6 * a:6 * a + 6].toarray()) for a in range(Kf)])
    nc = min(NC, max(Kf // 2, 1))
    agg = (Kf + nc - 1) // nc
    nc = (Kf + agg - 1) // agg
    a = np.arange(Kf)
    pos = np.clip((a + 0.5) / agg - 0.5, 0.0, nc - 1)
    lo = np.minimum(pos.astype(int), max(nc - 2, 0)); f = np.clip(pos - lo, 0.0, 1.0)
    pr, pc, pv = [], [], []
    for d in range(6):
        pr += [a * 6 + d, a * 6 + d]; pc += [lo * 6 + d, (lo + 1) * 6 + d]; pv += [1.0 - f, f]
    P = sp.csr_matrix((np.concatenate(pv), (np.concatenate(pr), np.concatenate(pc))), shape=(6 * Kf, 6 * nc))
    Ac = (P.T @ S @ P).toarray()
    Ainv = np.linalg.inv(Ac)


    def Minv(r):
        z = np.einsum("aij,aj->ai", Dinv, r.reshape(Kf, 6)).ravel()
        return z + P @ (Ainv @ (P.T @ r))


    def pcg_classic(tol, maxit=2000):
        x = np.zeros_like(b); r = b.copy(); z = Minv(r); pvec = z.copy(); rz = r @ z; bb = np.sqrt(b @ b)
        for it in range(1, maxit + 1):
            q = S @ pvec
            alpha = rz / (pvec @ q)                      # reduction 1
            x += alpha * pvec; r -= alpha * q
            z = Minv(r)
            rz_new = r @ z; rr = r @ r                  # reduction 2
            if np.sqrt(rr) <= tol * bb:
                return x, it
            pvec = z + (rz_new / rz) * pvec; rz = rz_new
        return x, maxit


    def pcg_single_reduction(tol, maxit=2000):
        """Chronopoulos-Gear: u = Minv r, w = S u; gamma = r.u, delta = w.u (and r.r) in ONE reduction."""
        x = np.zeros_like(b); r = b.copy(); bb = np.sqrt(b @ b)
        u = Minv(r); w = S @ u
        gamma = r @ u; delta = w @ u
        pvec = np.zeros_like(b); s = np.zeros_like(b)
        alpha, beta = gamma / delta, 0.0
        for it in range(1, maxit + 1):
            pvec = u + beta * pvec; s = w + beta * s     # s = S p by recurrence
            x += alpha * pvec; r -= alpha * s
            u = Minv(r); w = S @ u
            gamma_new = r @ u; delta = w @ u; rr = r @ r  # the one reduction of the iteration
            if np.sqrt(rr) <= tol * bb:
                return x, it
            beta = gamma_new / gamma
            alpha = gamma_new / (delta - beta * gamma_new / alpha)
            gamma = gamma_new
        return x, maxit


    xref = np.linalg.solve(Sd, b) if Sd is not None else None
    for tol in (1e-8, 1e-10):
        for nm, fn in (("classic (two reductions)", pcg_classic), ("single reduction (Chronopoulos-Gear)", pcg_single_reduction)):
            t0 = time.time(); x, it = fn(tol); dt = time.time() - t0
            true_res = np.linalg.norm(b - S @ x) / np.linalg.norm(b)
            err = np.linalg.norm(x - xref) / np.linalg.norm(xref) if xref is not None else float("nan")
            print(f"RESULT tol {tol:.0e}  {nm:40s} iterations {it:4d}  true residual {true_res:.3e}  error vs direct solve {err:.3e}  ({dt:.1f}s)", flush=True)



# LM divides lambda by up to 3 per accepted step: the later trials of a Global BA are the hard systems
for scale in (1.0, 1e-1, 1e-2, 1e-3, 1e-4):
    print(f'--- lambda = {scale:g} x initial', flush=True)
    run(lam0 * scale)
