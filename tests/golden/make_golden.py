"""Generate the golden fixtures under tests/golden/.

The reference ships no tests or golden vectors for this path ("parity unpinned"), and it cannot be compiled
here (Eigen / OpenCV headers / ROS absent).  These fixtures therefore pin the CPU oracle (oracle/, the restatement of the
reference algorithm) at the moment it agreed with the independent witnesses available in this container:

  * BA / LM      : tests/witness.py DenseLM (dense numpy normal equations + numpy solve, written separately from the oracle)
  * Lie algebra  : scipy.spatial.transform.Rotation
  * ORB stages   : cv2 4.13 (resize INTER_LINEAR, GaussianBlur 7x7 sigma 2 REFLECT_101, FastFeatureDetector 9_16 + NMS, fastAtan2)
  * Hamming      : numpy unpackbits
  * projection-guided matchers, DBoW2 transform : tests/witness_match.py (pure Python, brute-force window membership, dict containers)

Each block below asserts the agreement BEFORE writing, so a fixture can only be (re)generated from an oracle the witnesses
accept.  Run from the repo root:  python tests/golden/make_golden.py
The tests (tests/test_golden.py) then hold (a) the oracle and (b) the CUDA path to these files.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from ccm_slam_b200 import synth  # noqa: E402
from ccm_slam_b200.synth_images import make_image  # noqa: E402
from oracle import pyoracle as orc  # noqa: E402
import witness  # noqa: E402

HUBER_GBA = float(np.float32(np.sqrt(5.99)))    # (double)(float)sqrt(5.99): S/Optimizer.cpp:712
HUBER_LOCAL = float(np.float32(np.sqrt(5.991)))  # S/Optimizer.cpp:360
BA_KEYS = ("poses", "intr", "fixed", "points", "obs_kf", "obs_mp", "obs_uv", "obs_w")


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print("wrote %-22s %7.1f KB" % (name, os.path.getsize(path) / 1e3))


def ba_inputs(p):
    return {"in_" + k: getattr(p, k) for k in BA_KEYS}


def golden_known_answers():
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(7)
    ups = np.concatenate([
        rng.standard_normal((4, 6)) * 0.3,                      # generic
        rng.standard_normal((2, 6)) * 1e-7,                     # theta < 1e-5: the R = I + W + W^2, V = R branch (se3quat.h:223-257)
        np.array([[3.1, 0.01, -0.02, 0.5, -0.4, 0.3]]),         # near pi
        np.zeros((1, 6)),
    ])
    se3 = np.stack([orc.se3_exp(u) for u in ups])
    for u, qt in zip(ups, se3):
        th = np.linalg.norm(u[:3])
        if th >= 1e-5:  # the small-angle branch is deliberately not the exact exponential
            R = Rotation.from_rotvec(u[:3]).as_matrix()
            Ro = Rotation.from_quat(qt[:4]).as_matrix()
            assert np.abs(R - Ro).max() < 1e-12, "se3 exp vs scipy"
    s_ups = np.concatenate([rng.standard_normal((4, 7)) * 0.2, np.zeros((1, 7))])
    sim3 = np.stack([orc.sim3_exp(u) for u in s_ups])
    logs = np.stack([orc.sim3_log(s) for s in sim3])
    assert np.abs(logs - s_ups).max() < 1e-9, "sim3 log(exp(u)) round trip"
    d = HUBER_GBA
    d2 = float(np.float32(d * d))   # the vendored kernel keeps delta^2 in a float member (G/core/robust_kernel_impl.h:84)
    es = np.array([0.0, 1.0, d2 - 1e-9, d2, d2 + 1e-9, d * d, np.nextafter(max(d2, d * d), 10.0), 10.0, 1e4])
    hub = np.stack([orc.huber(e, d) for e in es])
    for e, (rho, w) in zip(es, hub[:, :2]):
        if e <= d2:
            assert rho == e and w == 1.0
        else:
            assert abs(rho - (2 * np.sqrt(e) * d - d2)) < 1e-12 and abs(w - d / np.sqrt(e)) < 1e-15
    # R -> q branches of Eigen's Quaterniond(Matrix3d): trace > 0 and the three "largest diagonal" cases
    Ts = []
    for rv in ([0.1, 0.2, -0.1], [3.0, 0.1, 0.1], [0.1, 3.0, 0.1], [0.1, 0.1, 3.0]):
        T = np.eye(4, dtype=np.float32)
        T[:3, :3] = Rotation.from_rotvec(rv).as_matrix().astype(np.float32)
        T[:3, 3] = np.float32([0.3, -1.2, 2.5])
        Ts.append(T)
    Ts = np.stack(Ts)
    qts = np.stack([orc.pose_from_Tcw_f32(T) for T in Ts])
    for T, qt in zip(Ts, qts):
        assert qt[3] >= 0 and abs(np.linalg.norm(qt[:4]) - 1) < 1e-15
        assert np.abs(Rotation.from_quat(qt[:4]).as_matrix() - T[:3, :3]).max() < 1e-6
    back = np.stack([orc.pose_to_Tcw_f32(qt) for qt in qts])
    zero, ones = np.zeros(32, np.uint8), np.full(32, 255, np.uint8)
    assert orc.descriptor_distance(zero, ones) == 256 and orc.descriptor_distance(ones, ones) == 0
    save("known_answers.npz", se3_upd=ups, se3_qt=se3, sim3_upd=s_ups, sim3=sim3, sim3_log=logs, huber_e=es, huber_delta=d,
         huber_out=hub, Tcw_f32=Ts, Tcw_qt=qts, Tcw_back_f32=back)


def golden_ba(name, cfg, iters):
    p = synth.make_config(cfg)
    ref = orc.ba_solve(p, iterations=iters, huber_delta=HUBER_GBA)
    w = witness.DenseLM(p, robust=True, delta=HUBER_GBA)
    wt = np.array(w.optimize(iters))  # rows: [it, lambda, chi2, rho, trials, lambda_after] like the oracle's trace
    n = ref["iters_done"]
    assert n == len(wt) and np.allclose(wt[:, 2], ref["trace"][:n, 2], rtol=1e-7), "oracle LM trace vs DenseLM witness"
    assert np.allclose(wt[:, 1], ref["trace"][:n, 1], rtol=1e-6) and np.array_equal(wt[:, 4], ref["trace"][:n, 4])
    save(name, **ba_inputs(p), iterations=iters, huber_delta=HUBER_GBA, iters_done=n, trials_total=ref["trials_total"],
         trace=ref["trace"][:n], poses=ref["poses"], points=ref["points"], chi2=ref["chi2"], depth_pos=ref["depth_pos"],
         witness_trace=wt)


def golden_local_ba():
    """LocalBundleAdjustmentClient protocol: optimize(5), level-1 + kernel drop for chi2 > 5.991 or depth <= 0, optimize(10)
    (S/Optimizer.cpp:536-587)."""
    p = synth.make_config("cfg2")
    r1 = orc.ba_solve(p, iterations=5, huber_delta=HUBER_LOCAL)
    out = (r1["chi2"] > 5.991) | (r1["depth_pos"] == 0)
    p2 = p.copy(); p2.poses = r1["poses"]; p2.points = r1["points"]; p2.edge_flags = out.astype(np.uint8) | 2
    r2 = orc.ba_solve(p2, iterations=10, huber_delta=HUBER_LOCAL, chi2_in=r1["chi2"])
    assert out.sum() > 0 and np.array_equal(r2["chi2"][out], r1["chi2"][out])
    save("ba_local_cfg2.npz", **ba_inputs(p), huber_delta=HUBER_LOCAL, r1_poses=r1["poses"], r1_points=r1["points"],
         r1_chi2=r1["chi2"], r1_depth_pos=r1["depth_pos"], r1_trace=r1["trace"][:r1["iters_done"]], flags=p2.edge_flags,
         r2_poses=r2["poses"], r2_points=r2["points"], r2_chi2=r2["chi2"], r2_depth_pos=r2["depth_pos"],
         r2_trace=r2["trace"][:r2["iters_done"]])


def golden_pgo():
    for K, fs in ((60, False), (60, True)):
        p = synth.make_pgo(K=K, fix_scale=fs)
        ref = orc.pgo_solve(p, iterations=20)
        ana = orc.pgo_solve(p, iterations=20, analytic_jac=True)  # witness: same LM with analytic instead of numeric Jacobians
        assert abs(ana["chi2_final"] - ref["chi2_final"]) <= 1e-4 * ref["chi2_final"] + 1e-12
        assert ref["chi2_final"] < (0.5 if fs else 0.1) * ref["chi2_initial"]  # the scale drift stays when the scale is fixed
        save("pgo_K%d_%s.npz" % (K, "fixscale" if fs else "free"), in_sim3=p.sim3, in_fixed=p.fixed, in_edge_i=p.edge_i,
             in_edge_j=p.edge_j, in_meas=p.meas, fix_scale=fs, iterations=20, sim3=ref["sim3"], chi2_initial=ref["chi2_initial"],
             chi2_final=ref["chi2_final"], trace=ref["trace"][:ref["iters_done"]])


def kp_arrays(k):
    return {f: np.asarray(k[f]) for f in ("x", "y", "size", "angle", "response", "octave")}


def golden_orb():
    import cv2
    for seed, w, h, store_image in ((3, 376, 240, True), (0, 752, 480, False)):
        img = make_image(seed, w, h)
        # stage witnesses: chained INTER_LINEAR pyramid, 7x7 sigma-2 blur, FAST 9_16 with NMS, all against cv2
        cur = img
        sc = np.float32(1.0)
        for lvl in range(1, 8):
            sc = np.float32(sc * np.float32(1.2))  # mvScaleFactor is a float chain (S/ORBextractor.cpp:585-592)
            inv = np.float32(1.0) / sc
            dw, dh = int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))
            nxt = orc.resize_linear_u8(cur, dw, dh)
            assert np.array_equal(nxt, cv2.resize(cur, (dw, dh), interpolation=cv2.INTER_LINEAR)), "resize vs cv2"
            cur = nxt
        assert np.array_equal(orc.gaussian_blur7(img), cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))
        det = cv2.FastFeatureDetector_create(20, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        cvk = [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in det.detect(img)]
        xy, score = orc.fast(img, 20)
        assert [(x, y, r) for (x, y), r in zip(xy.tolist(), score.tolist())] == cvk, "FAST (order, position, score) vs cv2"
        kps, desc = orc.orb_extract(img)
        extra = {"image": img} if store_image else {}
        save("orb_%dx%d_seed%d.npz" % (w, h, seed), seed=seed, width=w, height=h,
             image_sha256=np.frombuffer(hashlib.sha256(img.tobytes()).digest(), np.uint8), desc=desc, **kp_arrays(kps), **extra)


def golden_match():
    a = make_image(0); b = np.roll(make_image(0), (3, 5), axis=(0, 1))
    k1, d1 = orc.orb_extract(a); k2, d2 = orc.orb_extract(b)
    ref = np.unpackbits(d1[:64, None, :] ^ d2[None, :64, :], axis=2).sum(axis=2)
    got = np.array([[orc.descriptor_distance(x, y) for y in d2[:64]] for x in d1[:64]])
    assert np.array_equal(ref, got), "Hamming vs numpy"
    rng = np.random.default_rng(1)
    node = lambda d: (d[:, 0].astype(np.int64) * 7 + d[:, 1] // 64) % 97  # ~100 groups like DBoW2 at levelsup = 4
    n1, n2 = node(d1), node(d2)
    fv1, fv2 = orc.FeatureVector(n1), orc.FeatureVector(n2)
    has1 = (rng.random(len(d1)) < 0.7).astype(np.uint8); has2 = (rng.random(len(d2)) < 0.7).astype(np.uint8)
    out = {}
    for tag, nn, ori in (("a", 0.7, True), ("b", 0.9, False)):
        m, n = orc.match_bow_kf_frame(d1, has1, k1["angle"], fv1, d2, k2["angle"], fv2, nn, ori)
        assert n > 20
        out["kf_frame_" + tag] = m
        m, n = orc.match_bow_kf_kf(d1, has1, k1["angle"], fv1, d2, has2, k2["angle"], fv2, nn, ori)
        out["kf_kf_" + tag] = m
    fx, fy, cx, cy = [np.float32(v) for v in synth.EUROC_INTR]
    Kinv = np.linalg.inv(np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64))
    tx = np.array([[0, 0, 0], [0, 0, -1.0], [0, 1.0, 0]])
    F12 = (Kinv.T @ tx @ Kinv).astype(np.float32)
    sf = (1.2 ** np.arange(8)).astype(np.float32); ls2 = (sf * sf).astype(np.float32)
    v = lambda k, d, has, fv: dict(desc=d, has_mp=has, kp_xy=np.stack([k["x"], k["y"]], 1), octave=k["octave"],
                                   angle=k["angle"], fv=fv, intr=(fx, fy, cx, cy))
    for ori in (False, True):
        out["tri_ori%d" % ori] = orc.match_triangulation(v(k1, d1, has1, fv1), v(k2, d2, has2, fv2), F12, -5000.0, float(cy), ls2, sf, ori)
    save("match_shifted_pair.npz", d1=d1, d2=d2, node1=n1, node2=n2, has1=has1, has2=has2, angle1=k1["angle"], angle2=k2["angle"],
         octave1=k1["octave"], octave2=k2["octave"], xy1=np.stack([k1["x"], k1["y"]], 1), xy2=np.stack([k2["x"], k2["y"]], 1),
         F12=F12, ex=-5000.0, ey=float(cy), level_sigma2=ls2, scale_factors=sf, intr=np.float32([fx, fy, cx, cy]), **out)


GRID_KEYS = ("desc", "kp_xy", "octave", "angle")
QUERY_KEYS = ("valid", "uv", "radius", "level", "desc", "angle")


def pack_grid(prefix, g):
    out = {prefix + k: np.asarray(g[k]) for k in GRID_KEYS}
    out[prefix + "bounds"] = np.asarray(g["bounds"], np.float64); out[prefix + "cols_rows"] = np.array([g["cols"], g["rows"]], np.int32)
    return out


def pack_queries(prefix, q):
    return {prefix + k: np.asarray(q[k]) for k in QUERY_KEYS if k in q}


def golden_proj():
    """projection-guided matchers: written only where the pure-Python witness (tests/witness_match.py) agrees"""
    import witness_match as wm
    from ccm_slam_b200 import synth_match as sm
    g = sm.make_grid(n=400, seed=21); q = sm.make_queries(g, m=500, seed=22, th=4.0)
    rng = np.random.default_rng(23)
    has_obs = (rng.random(500) < 0.85).astype(np.uint8); blocked = (rng.random(400) < 0.2).astype(np.uint8)
    existing = np.where(rng.random(500) < 0.15, rng.integers(0, 400, 500), -1).astype(np.int32)
    out = {}
    m, n = orc.search_by_projection_track(g, q, has_obs, blocked, 0.8); w, wn = wm.search_track(g, q, has_obs, blocked, 0.8)
    assert n == wn and np.array_equal(m, w) and n > 100, "track vs witness"
    out["track_match"] = m
    for tag, reloc, od in (("last", False, 100), ("reloc", True, 64)):
        m, n = orc.search_by_projection_frame(g, q, has_obs, blocked, reloc, od, True); w, wn = wm.search_frame(g, q, has_obs, blocked, reloc, od, True)
        assert n == wn and np.array_equal(m, w) and n > 80, tag + " vs witness"
        out[tag + "_match"] = m
    b, m, n = orc.search_by_projection_sim3(g, q, blocked, existing); wb, w, wn = wm.search_sim3proj(g, q, blocked, existing)
    assert n == wn and np.array_equal(b, wb) and np.array_equal(m, w), "sim3 projection vs witness"
    out["sim3_best"] = b; out["sim3_match"] = m
    for tag, wts in (("fuse_chi2", sm.INV_LEVEL_SIGMA2), ("fuse_plain", None)):
        b, n = orc.fuse_search(g, q, wts); wb, wn = wm.fuse_search(g, q, wts)
        assert n == wn and np.array_equal(b, wb), tag + " vs witness"
        out[tag] = b
    # SearchBySim3: two keyframes sharing half of their features
    g1 = sm.make_grid(n=300, seed=24); g2 = sm.make_grid(n=320, seed=25)
    share = rng.permutation(300)[:150]
    g2["desc"][:150] = sm.flip_bits(g1["desc"][share], rng.integers(0, 30, 150), rng)
    g2["kp_xy"][:150] = g1["kp_xy"][share] + rng.normal(0, 1.5, (150, 2)).astype(np.float32); g2["octave"][:150] = g1["octave"][share]

    def queries(src_g, dst_g, ps, pd):
        mm = src_g["desc"].shape[0]
        uv = rng.uniform(0, 700, (mm, 2)).astype(np.float32); level = src_g["octave"].copy()
        uv[ps] = dst_g["kp_xy"][pd] + rng.normal(0, 1.0, (len(ps), 2)).astype(np.float32)
        return dict(valid=(rng.random(mm) < 0.8).astype(np.uint8), uv=uv, radius=(np.float32(7.5) * sm.SCALE_FACTORS[level]).astype(np.float32),
                    level=level, desc=src_g["desc"], angle=np.zeros(mm, np.float32))
    q12 = queries(g1, g2, share, np.arange(150)); q21 = queries(g2, g1, np.arange(150), share)
    m, n = orc.search_by_sim3(g1, g2, q12, q21); w, wn = wm.search_by_sim3(g1, g2, q12, q21)
    assert n == wn and np.array_equal(m, w) and n > 40, "SearchBySim3 vs witness"
    save("proj_matchers.npz", has_obs=has_obs, blocked=blocked, existing=existing, inv_level_sigma2=sm.INV_LEVEL_SIGMA2, mutual_match12=m,
         **pack_grid("g_", g), **pack_queries("q_", q), **pack_grid("g1_", g1), **pack_grid("g2_", g2), **pack_queries("q12_", q12),
         **pack_queries("q21_", q21), **out)


def golden_voc():
    """DBoW2 transform on a small synthetic vocabulary; witness = tests/witness_match.voc_transform"""
    import witness_match as wm
    from ccm_slam_b200 import synth_match as sm
    voc = sm.make_vocabulary(k=6, L=3, seed=31)
    feat = sm.make_voc_features(voc, n=300, seed=32)
    V = orc.Vocabulary(voc)
    out = {}
    for levelsup in (1, 2):
        r = V.transform(feat, levelsup); w = wm.voc_transform(voc, feat, levelsup)
        assert [(int(a), int(b), float(c)) for a, b, c in zip(r["word"], r["node"], r["weight"])] == w["per"], "descent vs witness"
        assert list(r["bow_id"]) == w["bow_id"] and list(r["bow_val"]) == w["bow_val"], "BowVector vs witness"
        assert list(r["fv_node_id"]) == list(w["fv"].keys())
        for k2, v in r.items():
            out["l%d_%s" % (levelsup, k2)] = v
    V.close()
    save("voc_k6_L3.npz", feat=feat, voc_k=voc["k"], voc_L=voc["L"], voc_scoring=voc["scoring"], voc_weighting=voc["weighting"],
         voc_parent=voc["parent"], voc_is_leaf=voc["is_leaf"], voc_desc=voc["desc"], voc_weight=voc["weight"], **out)


ALL = dict(known=golden_known_answers, ba_tiny=lambda: golden_ba("ba_tiny.npz", "tiny", 6), ba_small=lambda: golden_ba("ba_small.npz", "small", 8),
           local_ba=golden_local_ba, pgo=golden_pgo, orb=golden_orb, match=golden_match, proj=golden_proj, voc=golden_voc)

if __name__ == "__main__":
    orc.lib()
    for name in (sys.argv[1:] or list(ALL)):   # python tests/golden/make_golden.py [known ba_tiny ba_small local_ba pgo orb match proj voc]
        ALL[name]()
