"""Algorithm-level emulation of the row-distributed PCG (k_pcg2 with nranks > 1, csrc/pcg2.cuh; first written for its predecessor k_pcg_dist) with N virtual ranks in numpy: own-row SpMV, partial scalars summed in rank
order, z slices exchanged with iteration parity, p recomputed everywhere, partial restrictions summed, replicated coarse solve.
Compared with a serial PCG using the same preconditioner (M^-1 = blockdiag^-1 + P Ac^-1 P^T)."""
import numpy as np, scipy.sparse as sp
rng = np.random.default_rng(1)

def parents(a, agg, nc, prolong):
    if not prolong or nc < 2: return a // agg, a // agg, 1.0, 0.0
    pos = min(max((a + 0.5) / agg - 0.5, 0.0), float(nc - 1)); lo = min(int(pos), nc - 2); f = min(max(pos - lo, 0.0), 1.0)
    return lo, lo + 1, 1.0 - f, f

def make(n, BS, band):
    M = np.zeros((n * BS, n * BS))
    for a in range(n):
        for b in range(a, min(n, a + band)):
            B = rng.normal(size=(BS, BS)) * (1.0 if a != b else 3.0)
            M[a*BS:(a+1)*BS, b*BS:(b+1)*BS] = B; M[b*BS:(b+1)*BS, a*BS:(a+1)*BS] = B.T
    M = M @ M.T / (band * BS) + np.eye(n * BS) * 0.05
    return M

def serial(S, b, Dinv, P, Acinv, tol, maxit):
    Mv = lambda r: Dinv @ r + (P @ (Acinv @ (P.T @ r)) if P is not None else 0)
    x = np.zeros_like(b); r = b.copy(); z = Mv(r); p = z.copy(); rz = r @ z; bb = b @ b
    for it in range(1, maxit + 1):
        q = S @ p; a = rz / (p @ q); x += a * p; r -= a * q
        if r @ r <= tol * tol * bb: return x, it
        z = Mv(r); rz2 = r @ z; p = z + (rz2 / rz) * p; rz = rz2
    return x, maxit

def dist(S, b, Dinv_blocks, n, BS, N, agg, nc, prolong, Acinv, tol, maxit):
    nv = n * BS; nC = BS * nc if nc else 0
    cuts = [n * k // N for k in range(N + 1)]
    # per-rank state
    x = [np.zeros(nv) for _ in range(N)]; r = [np.zeros(nv) for _ in range(N)]; q = [np.zeros(nv) for _ in range(N)]
    p = [np.zeros((2, nv)) for _ in range(N)]
    zwin = [np.zeros((2, nv)) for _ in range(N)]; scal = [np.zeros((2, N, 4)) for _ in range(N)]; rcpart = [np.zeros((2, N, max(nC, 1))) for _ in range(N)]
    yc = [np.zeros(max(nC, 1)) for _ in range(N)]
    own = lambda k: slice(cuts[k] * BS, cuts[k + 1] * BS)
    for k in range(N): r[k][own(k)] = b[own(k)]
    coarse = nC > 0
    def coarse_correct(par):
        for k in range(N):
            rc = np.zeros(nC)
            for a in range(cuts[k], cuts[k + 1]):
                lo, hi, w0, w1 = parents(a, agg, nc, prolong)
                rc[lo*BS:(lo+1)*BS] += w0 * r[k][a*BS:(a+1)*BS]
                if w1: rc[hi*BS:(hi+1)*BS] += w1 * r[k][a*BS:(a+1)*BS]
            for dst in range(N): rcpart[dst][par, k, :nC] = rc
        for k in range(N):
            s = np.zeros(nC)
            for src in range(N): s += rcpart[k][par, src, :nC]
            yc[k][:nC] = Acinv @ s
    def precond(par, slot_rz, slot_rr, spar):
        for k in range(N):
            arz = arr = 0.0
            for a in range(cuts[k], cuts[k + 1]):
                rv = r[k][a*BS:(a+1)*BS]; zv = Dinv_blocks[a] @ rv
                if coarse:
                    lo, hi, w0, w1 = parents(a, agg, nc, prolong)
                    zv = zv + w0 * yc[k][lo*BS:(lo+1)*BS] + (w1 * yc[k][hi*BS:(hi+1)*BS] if w1 else 0)
                for dst in range(N): zwin[dst][par, a*BS:(a+1)*BS] = zv
                arz += rv @ zv; arr += rv @ rv
            for dst in range(N): scal[dst][spar, k, slot_rz] = arz; scal[dst][spar, k, slot_rr] = arr
    par = 0
    if coarse: coarse_correct(par)
    precond(par, 1, 2, 1)
    rz = [scal[k][1, :, 1].sum() for k in range(N)]; bb = [scal[k][1, :, 2].sum() for k in range(N)]
    assert len(set(rz)) == 1 and len(set(bb)) == 1
    rz, bb = rz[0], bb[0]; beta = 0.0; pc = 0
    for it in range(maxit):
        for k in range(N):
            z = zwin[k][par]; p[k][pc ^ 1] = beta * p[k][pc] + z          # all rows, every rank
            pq = 0.0
            for a in range(cuts[k], cuts[k + 1]):
                pv = beta * p[k][pc] + z                                    # formed on the fly from pold and z
                y = S[a*BS:(a+1)*BS, :] @ pv
                q[k][a*BS:(a+1)*BS] = y; pq += y @ (beta * p[k][pc][a*BS:(a+1)*BS] + z[a*BS:(a+1)*BS])
            for dst in range(N): scal[dst][par, k, 0] = pq
        pq = scal[0][par, :, 0].sum(); alpha = rz / pq
        for k in range(N):
            x[k][own(k)] += alpha * p[k][pc ^ 1][own(k)]; r[k][own(k)] -= alpha * q[k][own(k)]
        if coarse: coarse_correct(par)
        precond(par ^ 1, 1, 2, par)
        rz_new = scal[0][par, :, 1].sum(); rr = scal[0][par, :, 2].sum()
        assert all(np.array_equal(zwin[0][par ^ 1], zwin[k][par ^ 1]) for k in range(N))
        if rr <= tol * tol * bb:
            it += 1; break
        beta = rz_new / rz; rz = rz_new; pc ^= 1; par ^= 1
    xs = np.zeros(nv)
    for k in range(N): xs[own(k)] = x[k][own(k)]
    return xs, it

CASES = [(60, 6, 6, 2, 8, 0), (60, 6, 6, 4, 8, 1), (97, 6, 12, 8, 12, 1), (40, 7, 5, 3, 0, 0), (33, 6, 4, 8, 5, 1)]


def run_case(n, BS, band, N, nc_t, prolong):
    S = make(n, BS, band); b = rng.normal(size=n * BS)
    Db = [np.linalg.inv(S[a*BS:(a+1)*BS, a*BS:(a+1)*BS]) for a in range(n)]
    Dinv = sp.block_diag(Db).toarray()
    if nc_t:
        agg = -(-n // nc_t); nc = -(-n // agg)
        P = np.zeros((n * BS, nc * BS))
        for a in range(n):
            lo, hi, w0, w1 = parents(a, agg, nc, prolong)
            P[a*BS:(a+1)*BS, lo*BS:(lo+1)*BS] += w0 * np.eye(BS)
            if w1: P[a*BS:(a+1)*BS, hi*BS:(hi+1)*BS] += w1 * np.eye(BS)
        Acinv = np.linalg.inv(P.T @ S @ P)
    else:
        agg = 0; nc = 0; P = None; Acinv = None
    xs, its = serial(S, b, Dinv, P, Acinv, 1e-10, 500)
    xd, itd = dist(S, b, Db, n, BS, N, agg, nc, prolong, Acinv, 1e-10, 500)
    return dict(n=n, BS=BS, N=N, nc=nc, prolong=prolong, serial_its=its, dist_its=itd, rel_err=float(np.abs(xd - xs).max() / np.abs(xs).max()),
                residual=float(np.linalg.norm(S @ xd - b) / np.linalg.norm(b)))



if __name__ == "__main__":
    for c in CASES:
        r = run_case(*c)
        print("n={n} BS={BS} N={N} nc={nc} prolong={prolong}: serial {serial_its} its, distributed {dist_its} its, |x_d - x_s| / |x_s| = {rel_err:.2e}, residual {residual:.2e}".format(**r))
