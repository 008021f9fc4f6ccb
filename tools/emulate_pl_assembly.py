"""Host emulation of the piecewise-linear Galerkin assembly of k_pcg (csrc/pcg.cuh, prolong = 1): the running-sum state machine
(sums for coarse columns `cur` and `cur + 1`, shift when lo(b) advances, weighted flush into both row parents), lane by lane in
numpy, against scipy's P^T S P with an independently built P; also restriction and interpolation through coarse_parents().
The kernel path itself is the default since round 2 (parity suite green on B200, profiles/r2/prolong_parity.log)."""
import sys, numpy as np, scipy.sparse as sp

def parents(a, agg, nc):
    if nc < 2: return a // agg, a // agg, 1.0, 0.0
    pos = (a + 0.5) / agg - 0.5
    pos = min(max(pos, 0.0), float(nc - 1))
    lo = int(pos)
    if lo > nc - 2: lo = nc - 2
    f = min(max(pos - lo, 0.0), 1.0)
    return lo, lo + 1, 1.0 - f, f

def emulate(rowptr, col, val, n, BS, agg, nc):
    nC = BS * nc; BB = BS * BS
    A0 = np.zeros((nC, nC))
    for a in range(n):
        plo, phi, w0, w1 = parents(a, agg, nc)
        L = np.zeros(BB); H = np.zeros(BB); cur = -1
        def flush(J, T):
            if J < 0 or J >= nc: return
            for e in range(BB):
                r, c = e // BS, e % BS
                A0[plo * BS + r, J * BS + c] += w0 * T[e]
                if w1 != 0.0: A0[phi * BS + r, J * BS + c] += w1 * T[e]
        for j in range(rowptr[a], rowptr[a + 1]):
            blo, bhi, v0, v1 = parents(col[j], agg, nc)
            if blo != cur:
                if cur >= 0:
                    flush(cur, L)
                    if blo == cur + 1: L = H.copy()
                    else: flush(cur + 1, H); L = np.zeros(BB)
                H = np.zeros(BB); cur = blo
            s = val[j].reshape(-1)
            L += v0 * s; H += v1 * s
        if cur >= 0: flush(cur, L); flush(cur + 1, H)
    return A0

rng = np.random.default_rng(0)
for (n, BS, agg_target, band) in [(50, 6, 8, 5), (97, 6, 13, 40), (30, 7, 4, 3), (10, 6, 1, 2), (64, 6, 64, 10), (33, 6, 3, 33)]:
    # random SPD block-banded matrix with sorted columns, plus a few far off-band blocks (loop closures)
    M = np.zeros((n * BS, n * BS))
    for a in range(n):
        for b in range(a, min(n, a + band)):
            if a == b or rng.random() < 0.6:
                B = rng.normal(size=(BS, BS)); M[a*BS:(a+1)*BS, b*BS:(b+1)*BS] = B; M[b*BS:(b+1)*BS, a*BS:(a+1)*BS] = B.T
    for _ in range(3):
        a, b = sorted(rng.integers(0, n, 2))
        B = rng.normal(size=(BS, BS)); M[a*BS:(a+1)*BS, b*BS:(b+1)*BS] = B; M[b*BS:(b+1)*BS, a*BS:(a+1)*BS] = B.T
    M = M + M.T + np.eye(n * BS) * 50
    S = sp.bsr_matrix(M, blocksize=(BS, BS)); S.sort_indices()
    agg = -(-n // agg_target); nc = -(-n // agg)
    A0 = emulate(S.indptr, S.indices, S.data, n, BS, agg, nc)
    # reference: P from the study's prolongation (written independently), generalised to BS dofs
    k = np.arange(n)
    if nc >= 2:
        centre = (np.arange(nc) + 0.5) * agg - 0.5
        pos = np.clip((k - centre[0]) / agg, 0, nc - 1); lo = np.minimum(np.floor(pos).astype(int), nc - 2); f = np.clip(pos - lo, 0, 1)
        Pk = sp.csr_matrix((np.concatenate([1 - f, f]), (np.concatenate([k, k]), np.concatenate([lo, lo + 1]))), shape=(n, nc))
    else:
        Pk = sp.csr_matrix((np.ones(n), (k, k // agg)), shape=(n, nc))
    P = sp.kron(Pk, sp.identity(BS), format="csr")
    ref = (P.T @ S.tocsr() @ P).toarray()
    err = np.abs(A0 - ref).max() / np.abs(ref).max()
    # restriction / prolongation of a vector through parents()
    r = rng.normal(size=n * BS); rc = np.zeros(nc * BS); 
    for a in range(n):
        lo_, hi_, w0, w1 = parents(a, agg, nc)
        rc[lo_*BS:(lo_+1)*BS] += w0 * r[a*BS:(a+1)*BS]
        if w1: rc[hi_*BS:(hi_+1)*BS] += w1 * r[a*BS:(a+1)*BS]
    y = rng.normal(size=nc * BS); z = np.zeros(n * BS)
    for a in range(n):
        lo_, hi_, w0, w1 = parents(a, agg, nc)
        z[a*BS:(a+1)*BS] = w0 * y[lo_*BS:(lo_+1)*BS] + w1 * y[hi_*BS:(hi_+1)*BS]
    print(n, BS, "agg", agg, "nc", nc, "galerkin err", err, "restrict err", np.abs(rc - P.T @ r).max(), "prolong err", np.abs(z - P @ y).max(),
          "min eig Ac", np.linalg.eigvalsh(ref).min() > 0)
