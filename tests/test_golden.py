"""Golden-vector tests.  The fixtures under tests/golden/ were written by tests/golden/make_golden.py from the CPU oracle at a
moment it agreed with the independent witnesses (DenseLM, scipy, cv2, numpy); the reference itself ships none.

  * CPU (`-m "not gpu"`): the oracle still reproduces every fixture -> the checker cannot drift silently.
  * GPU (`-m gpu`): libccm_b200.so, through its C ABI, reproduces them: bit-exact for keypoints / descriptors / match indices
    / Hamming distances, 1e-4 relative on pose / landmark / Sim3 estimates (the north_star tolerance; measured ~1e-10).
"""
import hashlib
import os

import numpy as np
import pytest

from ccm_slam_b200 import synth
from ccm_slam_b200.synth_images import make_image

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KP_FIELDS = ("x", "y", "size", "angle", "response", "octave")


def load(name):
    return np.load(os.path.join(GOLD, name))


def ba_problem(g):
    return synth.BAProblem(poses=g["in_poses"], intr=g["in_intr"], fixed=g["in_fixed"], points=g["in_points"], obs_kf=g["in_obs_kf"],
                           obs_mp=g["in_obs_mp"], obs_uv=g["in_obs_uv"], obs_w=g["in_obs_w"])


def pgo_problem(g):
    return synth.PGOProblem(sim3=g["in_sim3"], fixed=g["in_fixed"], edge_i=g["in_edge_i"], edge_j=g["in_edge_j"], meas=g["in_meas"],
                            fix_scale=bool(g["fix_scale"]))


def orb_image(g):
    img = g["image"] if "image" in g.files else make_image(int(g["seed"]), int(g["width"]), int(g["height"]))
    assert hashlib.sha256(img.tobytes()).digest() == g["image_sha256"].tobytes(), "synthetic image generator drifted"
    return img


def close(a, b, tol):
    return np.abs(np.asarray(a) - np.asarray(b)).max() <= tol * max(1.0, np.abs(np.asarray(b)).max())


def match_inputs(g, FV):
    fx, fy, cx, cy = [np.float32(v) for v in g["intr"]]
    fv1, fv2 = FV(g["node1"]), FV(g["node2"])
    view = lambda i, fv: dict(desc=g["d%d" % i], has_mp=g["has%d" % i], kp_xy=g["xy%d" % i], octave=g["octave%d" % i],
                              angle=g["angle%d" % i], fv=fv, intr=(fx, fy, cx, cy))
    return fv1, fv2, view(1, fv1), view(2, fv2)


def unpack_grid(g, prefix):
    b = g[prefix + "bounds"]; cr = g[prefix + "cols_rows"]
    return dict(desc=g[prefix + "desc"], kp_xy=g[prefix + "kp_xy"], octave=g[prefix + "octave"], angle=g[prefix + "angle"],
                bounds=tuple(float(v) for v in b), cols=int(cr[0]), rows=int(cr[1]))


def unpack_queries(g, prefix):
    return {k: g[prefix + k] for k in ("valid", "uv", "radius", "level", "desc", "angle")}


def unpack_voc(g):
    return dict(k=int(g["voc_k"]), L=int(g["voc_L"]), scoring=int(g["voc_scoring"]), weighting=int(g["voc_weighting"]), parent=g["voc_parent"],
                is_leaf=g["voc_is_leaf"], desc=g["voc_desc"], weight=g["voc_weight"])


def check_proj_fixture(g, M):
    """M: object with the seven matcher calls (the oracle binding or the product's ORBmatcher adaptor) -> asserts the fixture outputs"""
    G, Q = unpack_grid(g, "g_"), unpack_queries(g, "q_")
    m, n = M.track(G, Q, g["has_obs"], g["blocked"], 0.8)
    assert np.array_equal(m, g["track_match"]) and n >= int((g["track_match"] >= 0).sum())   # n also counts overwritten assignments
    m, _ = M.frame(G, Q, g["has_obs"], g["blocked"], False, 100, True); assert np.array_equal(m, g["last_match"])
    m, _ = M.frame(G, Q, g["has_obs"], g["blocked"], True, 64, True); assert np.array_equal(m, g["reloc_match"])
    b, m, _ = M.sim3proj(G, Q, g["blocked"], g["existing"]); assert np.array_equal(b, g["sim3_best"]) and np.array_equal(m, g["sim3_match"])
    b, _ = M.fuse(G, Q, g["inv_level_sigma2"]); assert np.array_equal(b, g["fuse_chi2"])
    b, _ = M.fuse(G, Q, None); assert np.array_equal(b, g["fuse_plain"])
    m, n = M.mutual(unpack_grid(g, "g1_"), unpack_grid(g, "g2_"), unpack_queries(g, "q12_"), unpack_queries(g, "q21_"))
    assert np.array_equal(m, g["mutual_match12"]) and n == int((g["mutual_match12"] >= 0).sum())


# ----------------------------------------------------------------------------------------------- CPU: oracle vs fixtures
def test_oracle_reproduces_known_answers(oracle):
    g = load("known_answers.npz")
    assert np.array_equal(np.stack([oracle.se3_exp(u) for u in g["se3_upd"]]), g["se3_qt"])
    assert np.array_equal(np.stack([oracle.sim3_exp(u) for u in g["sim3_upd"]]), g["sim3"])
    assert np.array_equal(np.stack([oracle.sim3_log(s) for s in g["sim3"]]), g["sim3_log"])
    assert np.array_equal(np.stack([oracle.huber(e, float(g["huber_delta"])) for e in g["huber_e"]]), g["huber_out"])
    assert np.array_equal(np.stack([oracle.pose_from_Tcw_f32(T) for T in g["Tcw_f32"]]), g["Tcw_qt"])
    assert np.array_equal(np.stack([oracle.pose_to_Tcw_f32(q) for q in g["Tcw_qt"]]), g["Tcw_back_f32"])


@pytest.mark.parametrize("name", ["ba_tiny.npz", "ba_small.npz"])
def test_oracle_reproduces_ba_fixtures(oracle, name):
    g = load(name)
    r = oracle.ba_solve(ba_problem(g), iterations=int(g["iterations"]), huber_delta=float(g["huber_delta"]))
    assert r["iters_done"] == int(g["iters_done"]) and r["trials_total"] == int(g["trials_total"])
    assert np.allclose(r["trace"][:r["iters_done"]], g["trace"], rtol=1e-10, atol=0)
    assert np.allclose(r["poses"], g["poses"], rtol=0, atol=1e-12) and np.allclose(r["points"], g["points"], rtol=0, atol=1e-12)
    assert np.allclose(g["witness_trace"][:, 2], g["trace"][:, 2], rtol=1e-7)  # the witness agreement the fixture was written under


def test_oracle_reproduces_local_ba_fixture(oracle):
    g = load("ba_local_cfg2.npz")
    p = ba_problem(g); d = float(g["huber_delta"])
    r1 = oracle.ba_solve(p, iterations=5, huber_delta=d)
    assert np.allclose(r1["chi2"], g["r1_chi2"], rtol=1e-9) and np.array_equal(r1["depth_pos"], g["r1_depth_pos"])
    p2 = p.copy(); p2.poses = r1["poses"]; p2.points = r1["points"]; p2.edge_flags = g["flags"]
    r2 = oracle.ba_solve(p2, iterations=10, huber_delta=d, chi2_in=r1["chi2"])
    assert np.allclose(r2["trace"][:r2["iters_done"], 2], g["r2_trace"][:, 2], rtol=1e-9)
    assert np.allclose(r2["poses"], g["r2_poses"], atol=1e-10) and np.allclose(r2["points"], g["r2_points"], atol=1e-10)


@pytest.mark.parametrize("name", ["pgo_K60_free.npz", "pgo_K60_fixscale.npz"])
def test_oracle_reproduces_pgo_fixtures(oracle, name):
    g = load(name)
    r = oracle.pgo_solve(pgo_problem(g), iterations=int(g["iterations"]))
    assert np.allclose(r["sim3"], g["sim3"], atol=1e-9) and abs(r["chi2_final"] - float(g["chi2_final"])) <= 1e-9 * float(g["chi2_final"]) + 1e-15


@pytest.mark.parametrize("name", ["orb_376x240_seed3.npz", "orb_752x480_seed0.npz"])
def test_oracle_reproduces_orb_fixtures(oracle, name):
    g = load(name)
    kps, desc = oracle.orb_extract(orb_image(g))
    for f in KP_FIELDS:
        assert np.array_equal(kps[f], g[f]), f
    assert np.array_equal(desc, g["desc"])


def test_oracle_reproduces_match_fixture(oracle):
    g = load("match_shifted_pair.npz")
    fv1, fv2, v1, v2 = match_inputs(g, oracle.FeatureVector)
    for tag, nn, ori in (("a", 0.7, True), ("b", 0.9, False)):
        m, _ = oracle.match_bow_kf_frame(g["d1"], g["has1"], g["angle1"], fv1, g["d2"], g["angle2"], fv2, nn, ori)
        assert np.array_equal(m, g["kf_frame_" + tag])
        m, _ = oracle.match_bow_kf_kf(g["d1"], g["has1"], g["angle1"], fv1, g["d2"], g["has2"], g["angle2"], fv2, nn, ori)
        assert np.array_equal(m, g["kf_kf_" + tag])
    for ori in (False, True):
        m = oracle.match_triangulation(v1, v2, g["F12"], float(g["ex"]), float(g["ey"]), g["level_sigma2"], g["scale_factors"], ori)
        assert np.array_equal(m, g["tri_ori%d" % ori])


def test_host_selection_halves_reproduce_match_fixture(oracle):
    """the product's host halves of SearchByBoW x2 / SearchForTriangulation (ccm_select_*: everything after the distance matrix) on the
    CPU, fed a numpy Hamming matrix, against the fixture and the oracle — incl. a tie-storm variant where the visiting order decides"""
    from ccm_slam_b200.frontend import FeatureVector, ORBmatcher
    from ccm_slam_b200 import synth_match as sm
    g = dict(load("match_shifted_pair.npz"))
    for storm in (False, True):
        if storm:
            rng = np.random.default_rng(5)
            pats = np.concatenate([sm.flip_bits(rng.integers(0, 256, (1, 32), dtype=np.uint8), [int(rng.integers(8, 31))], rng) for _ in range(6)])
            g["d1"] = pats[rng.integers(0, 6, len(g["d1"]))]; g["d2"] = pats[rng.integers(0, 6, len(g["d2"]))]
        fv1, fv2, v1, v2 = match_inputs(g, FeatureVector)
        ofv1, ofv2, ov1, ov2 = match_inputs(g, oracle.FeatureVector)
        D = np.unpackbits(g["d1"][:, None, :] ^ g["d2"][None, :, :], axis=2).sum(axis=2).astype(np.uint16)
        for tag, nn, ori in (("a", 0.7, True), ("b", 0.9, False)):
            m = ORBmatcher(nn, ori)
            got, n = m.SearchByBoW_KF_Frame(g["d1"], g["has1"], g["angle1"], fv1, g["d2"], g["angle2"], fv2, D=D)
            ref, rn = oracle.match_bow_kf_frame(g["d1"], g["has1"], g["angle1"], ofv1, g["d2"], g["angle2"], ofv2, nn, ori)
            assert n == rn and np.array_equal(got, ref) and (storm or np.array_equal(got, g["kf_frame_" + tag]))
            got, n = m.SearchByBoW_KF_KF(g["d1"], g["has1"], g["angle1"], fv1, g["d2"], g["has2"], g["angle2"], fv2, D=D)
            ref, rn = oracle.match_bow_kf_kf(g["d1"], g["has1"], g["angle1"], ofv1, g["d2"], g["has2"], g["angle2"], ofv2, nn, ori)
            assert n == rn and np.array_equal(got, ref) and (storm or np.array_equal(got, g["kf_kf_" + tag]))
        for ori in (False, True):
            got = ORBmatcher(0.6, ori).SearchForTriangulation(v1, v2, g["F12"], float(g["ex"]), float(g["ey"]), g["level_sigma2"], g["scale_factors"], D=D)
            ref = oracle.match_triangulation(ov1, ov2, g["F12"], float(g["ex"]), float(g["ey"]), g["level_sigma2"], g["scale_factors"], ori)
            assert np.array_equal(got, ref) and (storm or np.array_equal(got, g["tri_ori%d" % ori]))


# ----------------------------------------------------------------------------------------------- GPU: C ABI vs fixtures
@pytest.fixture()
def gpu():
    from ccm_slam_b200 import api
    assert api.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    api.init(0)
    return api


def test_oracle_reproduces_proj_fixture(oracle):
    class M:
        track = staticmethod(oracle.search_by_projection_track)
        frame = staticmethod(oracle.search_by_projection_frame)
        sim3proj = staticmethod(oracle.search_by_projection_sim3)
        fuse = staticmethod(oracle.fuse_search)
        mutual = staticmethod(oracle.search_by_sim3)
    check_proj_fixture(load("proj_matchers.npz"), M)


def test_oracle_reproduces_voc_fixture(oracle):
    g = load("voc_k6_L3.npz")
    V = oracle.Vocabulary(unpack_voc(g))
    for levelsup in (1, 2):
        r = V.transform(g["feat"], levelsup)
        for k, v in r.items():
            assert np.array_equal(v, g["l%d_%s" % (levelsup, k)]), k
    V.close()


def test_library_host_halves_match_proj_and_voc_fixtures():
    """the product's selection / container code (no device involved) on the fixtures, distances from numpy"""
    from ccm_slam_b200.frontend import ORBmatcher, bow_assemble
    dist = lambda q, gr: np.unpackbits(q["desc"][:, None, :] ^ gr["desc"][None, :, :], axis=2).sum(axis=2).astype(np.uint16)

    class M:
        track = staticmethod(lambda G, Q, ho, bl, nn: ORBmatcher(nn).SearchByProjection_Track(G, Q, ho, bl, D=dist(Q, G)))
        frame = staticmethod(lambda G, Q, ho, bl, reloc, od, ori: ORBmatcher(0.9, ori).SearchByProjection_Frame(G, Q, ho, bl, reloc, od, D=dist(Q, G)))
        sim3proj = staticmethod(lambda G, Q, fm, ex: ORBmatcher().SearchByProjection_Sim3(G, Q, fm, ex, D=dist(Q, G)))
        fuse = staticmethod(lambda G, Q, w: ORBmatcher().Fuse(G, Q, w, D=dist(Q, G)))
        mutual = staticmethod(lambda G1, G2, Q12, Q21: ORBmatcher().SearchBySim3(G1, G2, Q12, Q21, D12=dist(Q12, G2), D21=dist(Q21, G1)))
    check_proj_fixture(load("proj_matchers.npz"), M)
    g = load("voc_k6_L3.npz")
    for levelsup in (1, 2):
        pre = "l%d_" % levelsup
        got = bow_assemble(int(g["voc_scoring"]), int(g["voc_weighting"]), g[pre + "word"], g[pre + "weight"], g[pre + "node"])
        for k in ("bow_id", "bow_val", "fv_node_id", "fv_node_ptr", "fv_feat"):
            assert np.array_equal(got[k], g[pre + k]), k


def test_library_pose_conversions_match_fixture():
    """Converter::toSE3Quat / toCvMat equivalents of the C ABI are host code: checked on the CPU box as well."""
    from ccm_slam_b200 import api
    g = load("known_answers.npz")
    assert np.array_equal(api.poses_from_Tcw_f32(g["Tcw_f32"]), g["Tcw_qt"])
    assert np.array_equal(api.poses_to_Tcw_f32(g["Tcw_qt"]), g["Tcw_back_f32"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ba_tiny.npz", "ba_small.npz"])
def test_gpu_ba_matches_fixture(gpu, name):
    g = load(name)
    r = gpu.ba_solve(ba_problem(g), iterations=int(g["iterations"]), huber_delta=float(g["huber_delta"]))
    assert r["iters_done"] == int(g["iters_done"]) and r["trials_total"] == int(g["trials_total"])
    n = r["iters_done"]
    assert np.allclose(r["trace"][:n, 2], g["trace"][:, 2], rtol=1e-7)           # robust chi2 after every LM iteration
    assert np.allclose(r["trace"][:n, 1], g["trace"][:, 1], rtol=1e-6)           # lambda used
    assert np.array_equal(r["trace"][:n, 4], g["trace"][:, 4])                   # trials per iteration
    assert close(r["poses"], g["poses"], 1e-4) and close(r["points"], g["points"], 1e-4)
    assert close(r["poses"], g["poses"], 1e-7) and close(r["points"], g["points"], 1e-7)   # what the f64 pipeline actually reaches
    assert np.array_equal(r["depth_pos"], g["depth_pos"]) and np.allclose(r["chi2"], g["chi2"], rtol=1e-5, atol=1e-9)


@pytest.mark.gpu
def test_gpu_local_ba_protocol_matches_fixture(gpu):
    g = load("ba_local_cfg2.npz")
    d = float(g["huber_delta"])
    h = gpu.BAHandle(ba_problem(g))
    r1 = h.optimize(iterations=5, huber_delta=d, want_edges=True)
    out = (r1["chi2"] > 5.991) | (r1["depth_pos"] == 0)
    assert np.array_equal(out.astype(np.uint8) | 2, g["flags"])
    h.set_edge_flags(g["flags"])
    r2 = h.optimize(iterations=10, huber_delta=d, want_edges=True, chi2_in=r1["chi2"])
    h.close()
    assert np.allclose(r2["trace"][:len(g["r2_trace"]), 2], g["r2_trace"][:, 2], rtol=1e-7)
    assert close(r2["poses"], g["r2_poses"], 1e-4) and close(r2["points"], g["r2_points"], 1e-4)
    erase = (r2["chi2"] > 5.991) | (r2["depth_pos"] == 0)
    assert np.array_equal(erase, (g["r2_chi2"] > 5.991) | (g["r2_depth_pos"] == 0))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["pgo_K60_free.npz", "pgo_K60_fixscale.npz"])
def test_gpu_pgo_matches_fixture(gpu, name):
    g = load(name)
    r = gpu.pgo_solve(pgo_problem(g), iterations=int(g["iterations"]))
    assert abs(r["chi2_initial"] - float(g["chi2_initial"])) <= 1e-9 * float(g["chi2_initial"])
    assert abs(r["chi2_final"] - float(g["chi2_final"])) <= 1e-4 * float(g["chi2_final"])
    assert close(r["sim3"], g["sim3"], 1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["orb_376x240_seed3.npz", "orb_752x480_seed0.npz"])
def test_gpu_orb_matches_fixture(gpu, name):
    from ccm_slam_b200.frontend import ORBextractor
    g = load(name)
    ex = ORBextractor(width=int(g["width"]), height=int(g["height"]))
    kps, desc = ex(orb_image(g))
    ex.close()
    for f in KP_FIELDS:
        assert np.array_equal(kps[f], g[f]), f
    assert np.array_equal(desc, g["desc"])


@pytest.mark.gpu
def test_gpu_matchers_match_fixture(gpu):
    from ccm_slam_b200.frontend import FeatureVector, ORBmatcher
    g = load("match_shifted_pair.npz")
    fv1, fv2, v1, v2 = match_inputs(g, FeatureVector)
    D = gpu.hamming_matrix(g["d1"], g["d2"])
    assert np.array_equal(D, np.unpackbits(g["d1"][:, None, :] ^ g["d2"][None, :, :], axis=2).sum(axis=2).astype(np.uint16))
    for tag, nn, ori in (("a", 0.7, True), ("b", 0.9, False)):
        m = ORBmatcher(nn, ori)
        got, _ = m.SearchByBoW_KF_Frame(g["d1"], g["has1"], g["angle1"], fv1, g["d2"], g["angle2"], fv2)
        assert np.array_equal(got, g["kf_frame_" + tag])
        got, _ = m.SearchByBoW_KF_KF(g["d1"], g["has1"], g["angle1"], fv1, g["d2"], g["has2"], g["angle2"], fv2)
        assert np.array_equal(got, g["kf_kf_" + tag])
    for ori in (False, True):
        got = ORBmatcher(0.6, ori).SearchForTriangulation(v1, v2, g["F12"], float(g["ex"]), float(g["ey"]), g["level_sigma2"], g["scale_factors"])
        assert np.array_equal(got, g["tri_ori%d" % ori])
