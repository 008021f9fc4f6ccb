"""Independent numpy/scipy witnesses used to pin the CPU oracle (the reference has no tests and cannot be built here).

Nothing in this file shares code with oracle/ or with the product: rotations go through scipy Rotation matrices,
the Jacobians are derived from the chain rule (not copied from g2o's closed forms), the linear system is the full
(un-Schur'ed) normal equation solved densely, and the LM schedule is restated from
G/core/optimization_algorithm_levenberg.cpp:61-164.
"""
import numpy as np
from scipy.spatial.transform import Rotation


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.0]])


def se3_exp(upd):
    """exp of (omega, upsilon) -> (R, t); includes g2o's small-angle quirk R = I + W + W^2, V = R (se3quat.h:237-243)."""
    om, up = np.asarray(upd[:3], float), np.asarray(upd[3:], float)
    th = np.linalg.norm(om)
    W = skew(om)
    if th < 1e-5:
        R = np.eye(3) + W + W @ W
        V = R
    else:
        R = Rotation.from_rotvec(om).as_matrix()
        V = np.eye(3) + (1 - np.cos(th)) / th**2 * W + (th - np.sin(th)) / th**3 * (W @ W)
    return R, V @ up


def qt_to_Rt(qt):
    return Rotation.from_quat(qt[:4]).as_matrix(), np.asarray(qt[4:7], float)


def residual(R, t, X, uv, intr):
    Xc = R @ X + t
    fx, fy, cx, cy = intr
    return np.array([uv[0] - (fx * Xc[0] / Xc[2] + cx), uv[1] - (fy * Xc[1] / Xc[2] + cy)]), Xc


def jacobians(R, Xc, intr):
    fx, fy, _, _ = intr
    x, y, z = Xc
    de_dXc = -np.array([[fx / z, 0, -fx * x / z**2], [0, fy / z, -fy * y / z**2]])
    J_point = de_dXc @ R
    J_pose = de_dXc @ np.hstack([-skew(Xc), np.eye(3)])  # Xc' = Xc + w x Xc + v
    return J_pose, J_point


def huber(e, delta):
    d2 = float(np.float32(delta * delta))   # the vendored RobustKernelHuber keeps delta^2 in a float (G/core/robust_kernel_impl.h:84)
    if e <= d2:
        return e, 1.0
    s = np.sqrt(e)
    return 2 * s * delta - d2, delta / s


class DenseLM:
    """Dense, un-Schur'ed Levenberg-Marquardt witness for small problems."""

    def __init__(self, p, robust=True, delta=np.sqrt(5.99)):
        self.p = p
        self.Rt = [qt_to_Rt(q) for q in p.poses]
        self.X = p.points.copy()
        self.robust, self.delta = robust, delta
        flags = p.edge_flags if p.edge_flags is not None else np.zeros(p.E, np.uint8)
        self.act = [e for e in range(p.E) if not (flags[e] & 1)]
        self.rob = [robust and not (flags[e] & 2) for e in range(p.E)]
        pose_has = np.zeros(p.K, bool); pt_has = np.zeros(p.P, bool)
        for e in self.act:
            pose_has[p.obs_kf[e]] = True; pt_has[p.obs_mp[e]] = True
        self.free_pose = [k for k in range(p.K) if pose_has[k] and not p.fixed[k]]
        self.free_pt = [j for j in range(p.P) if pt_has[j]]
        self.pi = {k: i for i, k in enumerate(self.free_pose)}
        self.li = {j: i for i, j in enumerate(self.free_pt)}
        self.n = 6 * len(self.free_pose) + 3 * len(self.free_pt)

    def chi2(self):
        tot = 0.0
        for e in self.act:
            p = self.p
            R, t = self.Rt[p.obs_kf[e]]
            r, _ = residual(R, t, self.X[p.obs_mp[e]], p.obs_uv[e].astype(float), p.intr[p.obs_kf[e]])
            c = float(p.obs_w[e]) * (r @ r)
            tot += huber(c, self.delta)[0] if self.rob[e] else c
        return tot

    def build(self):
        p = self.p
        H = np.zeros((self.n, self.n)); b = np.zeros(self.n)
        off = 6 * len(self.free_pose)
        for e in self.act:
            k, j = p.obs_kf[e], p.obs_mp[e]
            R, t = self.Rt[k]
            r, Xc = residual(R, t, self.X[j], p.obs_uv[e].astype(float), p.intr[k])
            Jp, Jl = jacobians(R, Xc, p.intr[k])
            w = float(p.obs_w[e])
            if self.rob[e]:
                w *= huber(w * (r @ r), self.delta)[1]
            sl = slice(off + 3 * self.li[j], off + 3 * self.li[j] + 3)
            H[sl, sl] += w * Jl.T @ Jl; b[sl] += -w * Jl.T @ r
            if k in self.pi:
                sp = slice(6 * self.pi[k], 6 * self.pi[k] + 6)
                H[sp, sp] += w * Jp.T @ Jp; b[sp] += -w * Jp.T @ r
                H[sp, sl] += w * Jp.T @ Jl; H[sl, sp] += w * Jl.T @ Jp
        return H, b

    def apply(self, x):
        off = 6 * len(self.free_pose)
        for k, i in self.pi.items():
            dR, dt = se3_exp(x[6 * i:6 * i + 6])
            R, t = self.Rt[k]
            self.Rt[k] = (dR @ R, dR @ t + dt)
        for j, i in self.li.items():
            self.X[j] = self.X[j] + x[off + 3 * i: off + 3 * i + 3]

    def optimize(self, iterations):
        trace = []
        lam, ni, nbad = -1.0, 2.0, 0
        for it in range(iterations):
            cur = self.chi2(); ini = cur
            H, b = self.build()
            if it == 0:
                lam = 1e-5 * np.abs(np.diag(H)).max(); ni = 2.0; nbad = 0
            q = 0
            while True:
                bak = (list(self.Rt), self.X.copy())
                lam_used = lam
                try:
                    x = np.linalg.solve(H + lam * np.eye(self.n), b); ok = True
                except np.linalg.LinAlgError:
                    x = np.zeros(self.n); ok = False
                self.apply(x)
                tmp = self.chi2() if ok else np.finfo(float).max
                rho = (cur - tmp) / (x @ (lam * x + b) + 1e-3)
                if rho > 0 and np.isfinite(tmp):
                    alpha = min(1 - (2 * rho - 1) ** 3, 2 / 3)
                    lam *= max(1 / 3, alpha); ni = 2.0; cur = tmp
                else:
                    lam *= ni; ni *= 2; self.Rt, self.X = bak[0], bak[1]
                q += 1
                if not (rho < 0 and q < 10):
                    break
            trace.append((it, lam_used, cur, rho, q, lam))
            if q == 10 or rho == 0:
                break
            nbad = nbad + 1 if (ini - cur) * 1e3 < ini else 0
            if nbad >= 3:
                break
        return np.array(trace)


# ---- single-vertex witnesses (PoseOptimizationClient, OptimizeSim3) ------------------------------------------------------
def _lm_single(n_dim, state, oplus, errors_fn, jac_fn, weights, active, robust, delta, iterations):
    """g2o's Levenberg loop (optimization_algorithm_levenberg.cpp:61-164) on one vertex with a dense normal equation.
    errors_fn(state) -> (E,2) residuals of ALL edges; jac_fn(state, e) -> (2, n_dim).  Returns (state, cached_errors)."""
    act = np.flatnonzero(active)
    err = np.zeros((len(weights), 2))
    if len(act) == 0:
        return state, err

    def refresh(st):
        allerr = errors_fn(st)
        err[act] = allerr[act]

    def chi():
        c = 0.0
        for e in act:
            v = weights[e] * (err[e] @ err[e])
            c += huber(v, delta)[0] if robust[e] else v
        return c

    lam, ni, nbad = -1.0, 2.0, 0
    for it in range(iterations):
        refresh(state)
        cur = chi(); ini = cur
        H = np.zeros((n_dim, n_dim)); b = np.zeros(n_dim)
        for e in act:
            J = jac_fn(state, e)
            w = weights[e]
            rw = huber(w * (err[e] @ err[e]), delta)[1] if robust[e] else 1.0
            H += J.T @ J * (w * rw)
            b += -J.T @ err[e] * (w * rw)
        if it == 0:
            lam = 1e-5 * np.abs(np.diag(H)).max(); ni = 2.0; nbad = 0
        q = 0
        while True:
            bak = state
            try:
                x = np.linalg.solve(H + lam * np.eye(n_dim), b)
                ok = bool(np.all(np.linalg.eigvalsh(H + lam * np.eye(n_dim)) > 0))
            except np.linalg.LinAlgError:
                x = np.zeros(n_dim); ok = False
            state = oplus(state, x)
            refresh(state)
            tmp = chi() if ok else np.finfo(float).max
            rho = (cur - tmp) / (x @ (lam * x + b) + 1e-3)
            if rho > 0 and np.isfinite(tmp):
                lam *= max(1 / 3, min(1 - (2 * rho - 1) ** 3, 2 / 3)); ni = 2.0; cur = tmp
            else:
                lam *= ni; ni *= 2; state = bak
            q += 1
            if not (rho < 0 and q < 10):
                break
        if q == 10 or rho == 0:
            break
        nbad = nbad + 1 if (ini - cur) * 1e3 < ini else 0
        if nbad >= 3:
            break
    return state, err


def pose_optimization(d):
    """PoseOptimizationClient (S/Optimizer.cpp:215-347) on make_pose_opt() data -> ((R,t), outlier mask, n_inliers)."""
    Xw = d["Xw"].astype(np.float64); uv = d["uv"].astype(np.float64); w = d["inv_sigma2"].astype(np.float64)
    intr = [float(v) for v in d["intr"]]
    n = len(w)
    if n < 3:
        return qt_to_Rt(d["Tcw0"]), np.zeros(n, bool), 0
    delta = float(np.float32(np.sqrt(5.991)))

    def errors(st):
        R, t = st
        Xc = Xw @ R.T + t
        return np.stack([uv[:, 0] - (intr[0] * Xc[:, 0] / Xc[:, 2] + intr[2]), uv[:, 1] - (intr[1] * Xc[:, 1] / Xc[:, 2] + intr[3])], 1)

    def jac(st, e):
        R, t = st
        return jacobians(R, R @ Xw[e] + t, intr)[0]

    def oplus(st, x):
        dR, dt = se3_exp(x)
        return dR @ st[0], dR @ st[1] + dt

    active = np.ones(n, bool); robust = np.ones(n, bool); outlier = np.zeros(n, bool)
    st0 = qt_to_Rt(d["Tcw0"])
    st = st0; nbad = 0
    for rnd in range(4):
        st, err = _lm_single(6, st0, oplus, errors, jac, w, active, robust, delta, 10)
        fresh = errors(st)
        err[outlier] = fresh[outlier]
        chi2 = (w * (err * err).sum(1)).astype(np.float32)
        outlier = chi2 > np.float32(5.991)
        active = ~outlier; nbad = int(outlier.sum())
        if rnd == 2:
            robust[:] = False
        if n < 10:
            break
    return st, outlier, n - nbad


def sim3_exp_matrix(u):
    """(sR, t) of exp([omega, upsilon, sigma]) through the 4x4 matrix exponential (independent of sim3.h's closed forms)."""
    from scipy.linalg import expm
    G = np.zeros((4, 4))
    G[:3, :3] = skew(u[:3]) + u[6] * np.eye(3)
    G[:3, 3] = u[3:6]
    M = expm(G)
    return M[:3, :3], M[:3, 3]


def sim3_optimization(d):
    """OptimizeSim3 (S/Optimizer.cpp:861-1056) on make_sim3_opt() data; state = (sR 3x3, t).  -> (state, inlier, nIn)."""
    P1 = d["P1c"].astype(np.float64); P2 = d["P2c"].astype(np.float64)
    uv1 = d["uv1"].astype(np.float64); uv2 = d["uv2"].astype(np.float64)
    n = len(P1)
    w = np.empty(2 * n); w[0::2] = d["w1"]; w[1::2] = d["w2"]
    K1 = [float(v) for v in d["K1"]]; K2 = [float(v) for v in d["K2"]]
    th2 = float(d["th2"]); delta = float(np.float32(np.sqrt(np.float32(th2))))
    S0 = d["S12_0"]
    st = (S0[7] * Rotation.from_quat(S0[:4]).as_matrix(), np.asarray(S0[4:7], float))

    def errors(s):
        A, t = s
        q1 = P2 @ A.T + t
        Ai = np.linalg.inv(A)
        q2 = (P1 - t) @ Ai.T
        e = np.empty((2 * n, 2))
        e[0::2, 0] = uv1[:, 0] - (K1[0] * q1[:, 0] / q1[:, 2] + K1[2]); e[0::2, 1] = uv1[:, 1] - (K1[1] * q1[:, 1] / q1[:, 2] + K1[3])
        e[1::2, 0] = uv2[:, 0] - (K2[0] * q2[:, 0] / q2[:, 2] + K2[2]); e[1::2, 1] = uv2[:, 1] - (K2[1] * q2[:, 1] / q2[:, 2] + K2[3])
        return e

    def oplus(s, x):
        x = np.array(x, float)
        if d["fix_scale"]:
            x[6] = 0
        dA, dt = sim3_exp_matrix(x)
        return dA @ s[0], dA @ s[1] + dt

    def jac(s, e):
        J = np.zeros((2, 7))
        for k in range(7):
            x = np.zeros(7); x[k] = 1e-9
            ep = errors(oplus(s, x))[e]; em = errors(oplus(s, -x))[e]
            J[:, k] = (ep - em) / 2e-9
        return J

    active = np.ones(2 * n, bool); robust = np.ones(2 * n, bool)
    st, err = _lm_single(7, st, oplus, errors, jac, w, active, robust, delta, 5)
    chi2 = w * (err * err).sum(1)
    inlier = ~((chi2[0::2] > th2) | (chi2[1::2] > th2))
    nbad = int((~inlier).sum())
    if n - nbad < 10:
        return None, inlier, 0
    active = np.repeat(inlier, 2)
    st, err2 = _lm_single(7, st, oplus, errors, jac, w, active, robust, delta, 10 if nbad > 0 else 5)
    chi2 = w * (err2 * err2).sum(1)
    keep = inlier & ~((chi2[0::2] > th2) | (chi2[1::2] > th2))
    return st, keep, int(keep.sum())
