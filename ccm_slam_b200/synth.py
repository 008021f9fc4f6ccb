"""Seeded synthetic BA / pose-graph problems with the shapes BASELINE.json names.

No EuRoC data and no ROS exist in this environment, so configs 1-4 are realised as synthetic problems of
EuRoC-like shape; config 5 is synthetic by definition.  Distributions
follow BASELINE.md: pose noise N(0, 0.02 rad / 0.05 m), point noise N(0, 0.05 m), pixel noise
N(0, sigma_octave) with the octave drawn proportionally to the per-level ORB quotas
217:181:151:126:105:87:73:60 (S/ORBextractor.cpp:604-615), invSigma2 = 1.2^(-2 octave)
(S/ORBextractor.cpp:584-600), 5 % gross outliers.  Intrinsics: cslam/conf/vi_euroc.yaml:9-12.

Pure numpy; shared by tests, bench.py and the oracle legs so that CPU and GPU paths see identical bytes.
"""
from __future__ import annotations

import dataclasses

import numpy as np

EUROC_INTR = (458.654, 457.296, 367.215, 248.375)
OCTAVE_QUOTAS = np.array([217, 181, 151, 126, 105, 87, 73, 60], dtype=np.float64)
SCALE_FACTOR = 1.2

# named configurations (BASELINE.json "configs"); (est.) shapes from BASELINE.md §3
CONFIGS = {
    "cfg1": dict(kind="local", n_local=15, n_fixed=10, P=2000, obs_per_point=6, seed=1),
    "cfg2": dict(kind="local", n_local=15, n_fixed=10, P=2000, obs_per_point=6, seed=1),
    "cfg3": dict(kind="global", K=400, P=25000, obs_per_point=6, window=24, n_agents=2, seed=2),
    "cfg4": dict(kind="global", K=800, P=50000, obs_per_point=6, window=24, n_agents=4, seed=3),
    "cfg5": dict(kind="global", K=10000, P=1000000, obs_per_point=20, window=40, n_agents=1, seed=4),
    # small shapes for parity tests
    "tiny": dict(kind="global", K=3, P=20, obs_per_point=3, window=3, n_agents=1, seed=7),
    "small": dict(kind="global", K=40, P=1500, obs_per_point=5, window=12, n_agents=2, seed=11),
}


@dataclasses.dataclass
class BAProblem:
    poses: np.ndarray      # (K,7) f64: qx qy qz qw tx ty tz  (Tcw)
    intr: np.ndarray       # (K,4) f64: fx fy cx cy (f32-representable)
    fixed: np.ndarray      # (K,) u8
    points: np.ndarray     # (P,3) f64 (f32-representable)
    obs_kf: np.ndarray     # (E,) i32
    obs_mp: np.ndarray     # (E,) i32
    obs_uv: np.ndarray     # (E,2) f32
    obs_w: np.ndarray      # (E,) f32  invSigma2
    edge_flags: np.ndarray | None = None  # (E,) u8
    gt_poses: np.ndarray | None = None
    gt_points: np.ndarray | None = None
    name: str = ""

    @property
    def K(self): return int(self.poses.shape[0])
    @property
    def P(self): return int(self.points.shape[0])
    @property
    def E(self): return int(self.obs_kf.shape[0])

    def copy(self):
        return dataclasses.replace(self, **{f.name: (getattr(self, f.name).copy() if isinstance(getattr(self, f.name), np.ndarray) else getattr(self, f.name)) for f in dataclasses.fields(self)})


def _quat_mul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def _quat_rot(q, v):
    u = np.cross(q[..., :3], v)
    u = u + u
    return v + q[..., 3:4] * u + np.cross(q[..., :3], u)


def _rotvec_to_quat(rv):
    th = np.linalg.norm(rv, axis=-1, keepdims=True)
    half = 0.5 * th
    k = np.where(th > 1e-12, np.sin(half) / np.maximum(th, 1e-300), 0.5)
    return np.concatenate([rv * k, np.cos(half)], axis=-1)


def _mat_to_quat(R):
    """Shepperd-style branchy conversion (vectorised); sign fixed to w>=0 and normalised like SE3Quat."""
    from scipy.spatial.transform import Rotation
    q = Rotation.from_matrix(R).as_quat()  # x y z w
    q = np.where(q[..., 3:4] < 0, -q, q)
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def _helix_cameras(K, rng, radius=2.0, dtheta=0.02, dz=0.002, phase=0.0, z0=0.0):
    """Ground-truth world->camera poses on a helix, optical axis pointing radially outward."""
    k = np.arange(K)
    th = phase + dtheta * k
    C = np.stack([radius * np.cos(th), radius * np.sin(th), z0 + dz * k], axis=-1)  # camera centres
    zc = np.stack([np.cos(th), np.sin(th), np.zeros(K)], axis=-1)                   # look direction
    up = np.tile(np.array([0.0, 0.0, 1.0]), (K, 1))
    xc = np.cross(up, zc); xc /= np.linalg.norm(xc, axis=-1, keepdims=True)
    yc = np.cross(zc, xc)
    Rcw = np.stack([xc, yc, zc], axis=1)  # rows are camera axes in world -> Xc = Rcw (X - C)
    tcw = -np.einsum("kij,kj->ki", Rcw, C)
    return Rcw, tcw, C, zc


def make_global_ba(K, P, obs_per_point, window, n_agents=1, seed=0, outlier_frac=0.05,
                   pose_noise=(0.02, 0.05), point_noise=0.05, shared_frac=0.10, name="") -> BAProblem:
    rng = np.random.default_rng(seed)
    Ka = K // n_agents
    Rs, ts, Cs, Zs = [], [], [], []
    # the covisibility window spans 0.4 rad of the circle so that co-observers stay inside the 752x480 FOV
    dtheta = 0.4 / max(window, obs_per_point)
    for a in range(n_agents):
        n = Ka if a < n_agents - 1 else K - Ka * (n_agents - 1)
        R, t, C, z = _helix_cameras(n, rng, phase=0.0 + 0.3 * a, z0=0.15 * a, dtheta=dtheta)
        Rs.append(R); ts.append(t); Cs.append(C); Zs.append(z)
    Rcw = np.concatenate(Rs); tcw = np.concatenate(ts); C = np.concatenate(Cs); zdir = np.concatenate(Zs)
    agent_of = np.concatenate([np.full(len(r), a) for a, r in enumerate(Rs)])
    agent_start = np.concatenate([[0], np.cumsum([len(r) for r in Rs])])

    # each landmark: centre keyframe c (spread uniformly), placed in front of c at depth U[2,10]
    centre = np.sort(rng.integers(0, K, size=P))
    depth = rng.uniform(2.0, 10.0, size=P)
    lat = rng.uniform(-0.25, 0.25, size=(P, 2)) * depth[:, None]  # lateral offsets in the camera frame
    Xc = np.stack([lat[:, 0], lat[:, 1], depth], axis=-1)
    gt_pts = np.einsum("pji,pj->pi", Rcw[centre], Xc - tcw[centre])  # R^T (Xc - t)

    n_obs = obs_per_point
    w = max(window, n_obs)
    a_of_c = agent_of[centre]
    lo = np.maximum(agent_start[a_of_c], centre - w // 2)
    hi = np.minimum(agent_start[a_of_c + 1], lo + w)
    lo = np.maximum(agent_start[a_of_c], hi - w)
    span = hi - lo
    # choose n_obs distinct offsets in [0, span) per landmark
    keys = rng.random((P, w))
    keys[np.arange(w)[None, :] >= span[:, None]] = 2.0
    sel = np.sort(np.argpartition(keys, n_obs - 1, axis=1)[:, :n_obs], axis=1)
    valid = sel < span[:, None]
    kf = lo[:, None] + sel
    # inter-agent sharing: a fraction of landmarks gets half of its observers moved to another agent's
    # keyframes at the same trajectory phase (map-merge overlap)
    if n_agents > 1 and shared_frac > 0:
        shared = rng.random(P) < shared_frac
        other = (a_of_c + 1 + rng.integers(0, n_agents - 1, size=P)) % n_agents
        rel = centre - agent_start[a_of_c] + np.rint(0.3 * (a_of_c - other) / dtheta).astype(np.int64)  # same phase
        oc = agent_start[other] + np.clip(rel, 0, (agent_start[other + 1] - agent_start[other]) - 1)
        half = n_obs // 2
        okf = oc[:, None] + np.arange(-(half // 2), half - half // 2)[None, :]
        okf = np.clip(okf, agent_start[other][:, None], agent_start[other + 1][:, None] - 1)
        kf[shared, :half] = okf[shared]
        kf.sort(axis=1)
        dup = np.zeros_like(valid)
        dup[:, 1:] = kf[:, 1:] == kf[:, :-1]
        valid &= ~dup
    mp = np.broadcast_to(np.arange(P)[:, None], kf.shape)
    kf = kf[valid].astype(np.int32); mp = mp[valid].astype(np.int32)

    # ground-truth projections
    fx, fy, cx, cy = EUROC_INTR
    Xcam = np.einsum("eij,ej->ei", Rcw[kf], gt_pts[mp]) + tcw[kf]
    good = (Xcam[:, 2] > 0.5) & (np.abs(fx * Xcam[:, 0] / Xcam[:, 2]) < 1.1 * cx) & (np.abs(fy * Xcam[:, 1] / Xcam[:, 2]) < 1.1 * cy)
    kf, mp, Xcam = kf[good], mp[good], Xcam[good]
    E = kf.shape[0]
    octave = rng.choice(8, size=E, p=OCTAVE_QUOTAS / OCTAVE_QUOTAS.sum())
    sigma = SCALE_FACTOR ** octave
    uv = np.stack([fx * Xcam[:, 0] / Xcam[:, 2] + cx, fy * Xcam[:, 1] / Xcam[:, 2] + cy], axis=-1)
    uv += rng.normal(size=(E, 2)) * sigma[:, None]
    out = rng.random(E) < outlier_frac
    ang = rng.uniform(0, 2 * np.pi, size=E)
    mag = rng.uniform(10.0, 50.0, size=E)
    uv[out] += (np.stack([np.cos(ang), np.sin(ang)], -1) * mag[:, None])[out]
    inv_sigma2 = (1.0 / (SCALE_FACTOR ** (2 * octave))).astype(np.float32)

    # perturbed initial estimates
    q_gt = _mat_to_quat(Rcw)
    drot = _rotvec_to_quat(rng.normal(size=(K, 3)) * pose_noise[0])
    dt = rng.normal(size=(K, 3)) * pose_noise[1]
    q0 = _quat_mul(drot, q_gt)
    t0 = _quat_rot(drot, tcw) + dt
    fixed = np.zeros(K, np.uint8); fixed[0] = 1
    q0[0] = q_gt[0]; t0[0] = tcw[0]
    q0 = np.where(q0[:, 3:4] < 0, -q0, q0)
    q0 /= np.linalg.norm(q0, axis=-1, keepdims=True)
    pts0 = (gt_pts + rng.normal(size=(P, 3)) * point_noise).astype(np.float32).astype(np.float64)
    t0 = t0.astype(np.float32).astype(np.float64)  # translations arrive as f32 (S/Converter.cc:48)

    intr = np.tile(np.array(EUROC_INTR, np.float32).astype(np.float64), (K, 1))
    return BAProblem(poses=np.ascontiguousarray(np.concatenate([q0, t0], -1)), intr=intr, fixed=fixed,
                     points=np.ascontiguousarray(pts0), obs_kf=np.ascontiguousarray(kf),
                     obs_mp=np.ascontiguousarray(mp), obs_uv=np.ascontiguousarray(uv.astype(np.float32)),
                     obs_w=np.ascontiguousarray(inv_sigma2),
                     gt_poses=np.concatenate([q_gt, tcw], -1), gt_points=gt_pts, name=name)


def make_local_ba(n_local=15, n_fixed=10, P=2000, obs_per_point=6, seed=1, name="") -> BAProblem:
    """LocalBundleAdjustmentClient-shaped window: the newest n_local KFs are free, the n_fixed older ones that
    co-observe the local points are fixed (S/Optimizer.cpp:351-404)."""
    K = n_local + n_fixed
    p = make_global_ba(K, P, obs_per_point, window=K, n_agents=1, seed=seed, name=name)
    p.fixed[:] = 0
    p.fixed[:n_fixed] = 1
    # fixed keyframes are consistent with the map in the reference (they were optimised before): keep them at GT
    gt = p.gt_poses[:n_fixed].copy()
    gt[:, 4:] = gt[:, 4:].astype(np.float32)
    p.poses[:n_fixed] = gt
    return p


def make_config(name: str, **over) -> BAProblem:
    cfg = dict(CONFIGS[name]); cfg.update(over)
    kind = cfg.pop("kind")
    if kind == "local":
        return make_local_ba(name=name, **cfg)
    return make_global_ba(name=name, **cfg)


@dataclasses.dataclass
class PGOProblem:
    sim3: np.ndarray      # (K,8) qx qy qz qw tx ty tz s
    fixed: np.ndarray     # (K,) u8
    edge_i: np.ndarray    # (E,) i32
    edge_j: np.ndarray    # (E,) i32
    meas: np.ndarray      # (E,8) Sji
    fix_scale: bool = False
    gt: np.ndarray | None = None


def _sim3_mul(a, b):
    q = _quat_mul(a[..., :4], b[..., :4])
    t = a[..., 7:8] * _quat_rot(a[..., :4], b[..., 4:7]) + a[..., 4:7]
    return np.concatenate([q, t, a[..., 7:8] * b[..., 7:8]], -1)


def _sim3_inv(a):
    qc = a[..., :4] * np.array([-1, -1, -1, 1.0])
    t = _quat_rot(qc, (-1.0 / a[..., 7:8]) * a[..., 4:7])
    return np.concatenate([qc, t, 1.0 / a[..., 7:8]], -1)


def make_pgo(K=200, n_loop=6, n_covis=3, seed=5, drift=(0.002, 0.01, 0.002), fix_scale=False) -> PGOProblem:
    """Essential-graph-shaped Sim3 pose graph: spanning-tree chain + covisibility edges to the previous
    n_covis keyframes + a few loop edges (S/Optimizer.cpp:1389-1508).  Measurements are built like the
    reference does (Sji = Sjw * Swi from the *current*, drifted estimate) except for the loop edges which
    come from ground truth - that is what creates the error the optimisation distributes."""
    rng = np.random.default_rng(seed)
    Rcw, tcw, _, _ = _helix_cameras(K, rng, dtheta=2 * np.pi / K * 1.0, dz=0.0)
    q = _mat_to_quat(Rcw)
    gt = np.concatenate([q, tcw, np.ones((K, 1))], -1)
    # accumulate drift along the chain
    est = gt.copy()
    acc = np.array([0, 0, 0, 1.0, 0, 0, 0, 1.0])
    for k in range(1, K):
        d = np.concatenate([_rotvec_to_quat(rng.normal(size=3) * drift[0]), rng.normal(size=3) * drift[1],
                            [np.exp(rng.normal() * drift[2] * (0 if fix_scale else 1))]])
        acc = _sim3_mul(d, acc)
        est[k] = _sim3_mul(acc, gt[k])
    ei, ej, meas = [], [], []
    def add(i, j, src):
        ei.append(i); ej.append(j)
        meas.append(_sim3_mul(src[j], _sim3_inv(src[i])))
    for k in range(1, K):
        add(k, k - 1, est)
        for c in range(2, n_covis + 2):
            if k - c >= 0: add(k, k - c, est)
    for l in range(n_loop):
        i = K - 1 - l * 2
        j = l * 2
        add(i, j, gt)
    fixed = np.zeros(K, np.uint8); fixed[0] = 1
    return PGOProblem(sim3=est, fixed=fixed, edge_i=np.array(ei, np.int32), edge_j=np.array(ej, np.int32),
                      meas=np.ascontiguousarray(np.array(meas)), fix_scale=fix_scale, gt=gt)


# ---- single-vertex problems: PoseOptimizationClient (one frame against its map points), OptimizeSim3 (two keyframes) ----
def make_pose_opt(n=300, seed=11, outlier_frac=0.15, noise_px=0.8, pose_noise=(0.03, 0.08)):
    """One frame: n map points in front of a camera, pixel noise scaled by the octave, a share of gross outliers, and a
    perturbed initial pose.  Returns dict(Tcw0, Xw, uv, inv_sigma2, intr, Tcw_gt)."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = [np.float32(v) for v in EUROC_INTR]
    q_gt = _rotvec_to_quat(rng.standard_normal(3) * 0.3); t_gt = rng.standard_normal(3) * 0.5
    Xc = np.stack([rng.uniform(-2.0, 2.0, n), rng.uniform(-1.2, 1.2, n), rng.uniform(2.0, 8.0, n)], 1)
    qc = np.array([-q_gt[0], -q_gt[1], -q_gt[2], q_gt[3]])
    Xw = np.stack([_quat_rot(qc, x - t_gt) for x in Xc]).astype(np.float32)
    octave = rng.integers(0, 8, n)
    sigma = (1.2 ** octave)
    Xcf = np.stack([_quat_rot(q_gt, x.astype(np.float64)) + t_gt for x in Xw])
    uv = np.stack([fx * Xcf[:, 0] / Xcf[:, 2] + cx, fy * Xcf[:, 1] / Xcf[:, 2] + cy], 1) + rng.standard_normal((n, 2)) * noise_px * sigma[:, None]
    bad = rng.random(n) < outlier_frac
    uv[bad] += rng.uniform(-60, 60, (int(bad.sum()), 2))
    q0 = _quat_mul(_rotvec_to_quat(rng.standard_normal(3) * pose_noise[0]), q_gt)
    t0 = t_gt + rng.standard_normal(3) * pose_noise[1]
    return dict(Tcw0=np.concatenate([q0 / np.linalg.norm(q0), t0]), Xw=Xw, uv=uv.astype(np.float32),
                inv_sigma2=(1.0 / (sigma * sigma)).astype(np.float32), intr=(fx, fy, cx, cy), Tcw_gt=np.concatenate([q_gt, t_gt]))


def make_sim3_opt(n=120, seed=12, outlier_frac=0.2, noise_px=0.7, scale=1.15, fix_scale=False):
    """Two keyframes seeing the same n points, each with its own copy of the point in its own camera frame (map 2 is scaled
    against map 1).  Returns dict(S12_0, P1c, P2c, uv1, uv2, w1, w2, K1, K2, th2, fix_scale, S12_gt)."""
    rng = np.random.default_rng(seed)
    K = tuple(np.float32(v) for v in EUROC_INTR)
    s = 1.0 if fix_scale else scale
    q12 = _rotvec_to_quat(rng.standard_normal(3) * 0.2); t12 = rng.standard_normal(3) * 0.4
    P2c = np.stack([rng.uniform(-1.5, 1.5, n), rng.uniform(-1.0, 1.0, n), rng.uniform(2.5, 7.0, n)], 1)
    P1c = np.stack([s * _quat_rot(q12, x) + t12 for x in P2c])
    octave1, octave2 = rng.integers(0, 8, n), rng.integers(0, 8, n)
    s1, s2 = 1.2 ** octave1, 1.2 ** octave2
    proj = lambda P: np.stack([K[0] * P[:, 0] / P[:, 2] + K[2], K[1] * P[:, 1] / P[:, 2] + K[3]], 1)
    uv1 = proj(P1c) + rng.standard_normal((n, 2)) * noise_px * s1[:, None]
    uv2 = proj(P2c) + rng.standard_normal((n, 2)) * noise_px * s2[:, None]
    bad = rng.random(n) < outlier_frac
    uv1[bad] += rng.uniform(-40, 40, (int(bad.sum()), 2))
    P1c_n = P1c + rng.standard_normal((n, 3)) * 0.01; P2c_n = P2c + rng.standard_normal((n, 3)) * 0.01
    q0 = _quat_mul(_rotvec_to_quat(rng.standard_normal(3) * 0.02), q12)
    S0 = np.concatenate([q0 / np.linalg.norm(q0), t12 + rng.standard_normal(3) * 0.05, [s * (1.0 if fix_scale else 1.03)]])
    return dict(S12_0=S0, P1c=P1c_n.astype(np.float32), P2c=P2c_n.astype(np.float32), uv1=uv1.astype(np.float32), uv2=uv2.astype(np.float32),
                w1=(1.0 / (s1 * s1)).astype(np.float32), w2=(1.0 / (s2 * s2)).astype(np.float32), K1=K, K2=K, th2=np.float32(10.0),
                fix_scale=fix_scale, S12_gt=np.concatenate([q12, t12, [s]]))


def make_map_update(K=200, P=5000, seed=0, new_kf_frac=0.15, outside_frac=0.03, n_origins=2, chain=0.7):
    """A map as Map::RunGBA (S/Map.cpp:1441-1570) finds it when MapFusionGBA returns: a spanning forest over K keyframes rooted at
    n_origins origins (chain-like: with probability `chain` a keyframe hangs under its predecessor), f32 poses, a BA result for the
    keyframes that existed when the BA started and none for those added meanwhile (new_kf_frac, never an origin), a few keyframes
    outside the tree (outside_frac, some of them holding a BA result), points with and without a BA result, bad points, points
    without a reference keyframe."""
    rng = np.random.default_rng(seed)

    def se3(n, rot=0.6, trans=8.0):
        w = rng.normal(0, rot, (n, 3)); th = np.linalg.norm(w, axis=1, keepdims=True); k = w / np.maximum(th, 1e-12)
        Kx = np.zeros((n, 3, 3)); Kx[:, 0, 1] = -k[:, 2]; Kx[:, 0, 2] = k[:, 1]; Kx[:, 1, 0] = k[:, 2]; Kx[:, 1, 2] = -k[:, 0]; Kx[:, 2, 0] = -k[:, 1]; Kx[:, 2, 1] = k[:, 0]
        R = np.eye(3) + np.sin(th)[:, :, None] * Kx + (1 - np.cos(th))[:, :, None] * (Kx @ Kx)
        T = np.tile(np.eye(4), (n, 1, 1)); T[:, :3, :3] = R; T[:, :3, 3] = rng.normal(0, trans, (n, 3))
        return T
    parent = np.full(K, -2, np.int32)
    n_origins = min(n_origins, K)
    outside = np.zeros(K, bool)
    if K > n_origins:
        outside[n_origins:] = rng.random(K - n_origins) < outside_frac
    last_in = []
    for k in range(K):
        if k < n_origins:
            parent[k] = -1
        elif not outside[k]:
            parent[k] = last_in[-1] if rng.random() < chain else last_in[int(rng.integers(0, len(last_in)))]
        if not outside[k]:
            last_in.append(k)
    optimized = rng.random(K) >= new_kf_frac
    optimized[:n_origins] = True
    Tcw = se3(K).astype(np.float32)
    corr = se3(K, rot=0.02, trans=0.15)
    gba = (corr @ Tcw.astype(np.float64)).astype(np.float32)
    gba[~optimized] = np.float32(np.nan)                      # mTcwGBA is an empty Mat there; the update must never read it
    state = rng.choice(np.array([0, 1, 2], np.uint8), P, p=[0.05, 0.8, 0.15]) if P else np.zeros(0, np.uint8)
    ref = rng.integers(0, K, P).astype(np.int32) if K and P else np.full(P, -1, np.int32)
    if P:
        ref[rng.random(P) < 0.03] = -1
    pos = rng.normal(0, 12.0, (P, 3)).astype(np.float32)
    pos_gba = (pos + rng.normal(0, 0.05, (P, 3))).astype(np.float32)
    pos_gba[state != 1] = np.float32(np.nan)
    return dict(kf_parent=parent, kf_optimized=optimized.astype(np.uint8), kf_Tcw=Tcw, kf_TcwGBA=gba, mp_state=state, mp_ref=ref, mp_pos=pos,
                mp_pos_gba=pos_gba)
