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
    d2 = delta * delta
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
