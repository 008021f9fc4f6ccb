"""GPU parity tests of the ORB front end, the matchers and the Sim3 pose graph: C ABI vs CPU oracle, bit-exact for the
integer/byte/index work, 1e-4 relative for the Sim3 estimates."""
import numpy as np
import pytest

from ccm_slam_b200 import api, synth
from ccm_slam_b200.frontend import FeatureVector, ORBextractor, ORBmatcher
from ccm_slam_b200.synth_images import make_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _dev():
    assert api.device_count() > 0
    api.init(0)


@pytest.mark.parametrize("seed,w,h", [(0, 752, 480), (1, 752, 480), (2, 640, 480), (3, 376, 240)])
def test_orb_extract_bit_exact(oracle, seed, w, h):
    img = make_image(seed, w, h)
    ex = ORBextractor(width=w, height=h)
    kps, desc = ex(img)
    rk, rd = oracle.orb_extract(img)
    assert len(kps) == len(rk) > 100
    for f in ("x", "y", "size", "angle", "response", "octave"):
        assert np.array_equal(kps[f], rk[f]), f
    assert np.array_equal(desc, rd)
    # mvImagePyramid levels
    cur = img
    for l in range(1, 8):
        lv = ex.image_pyramid(l)
        cur = oracle.resize_linear_u8(cur, lv.shape[1], lv.shape[0])
        assert np.array_equal(lv, cur)
    # a second frame through the same handle
    img2 = make_image(seed + 10, w, h)
    k2, d2 = ex(img2); r2k, r2d = oracle.orb_extract(img2)
    assert np.array_equal(k2["x"], r2k["x"]) and np.array_equal(d2, r2d)
    ex.close()


def test_orb_edge_cases(oracle):
    ex = ORBextractor()
    flat = np.full((480, 752), 127, np.uint8)
    kps, desc = ex(flat)
    assert len(kps) == 0 and desc.shape == (0, 32)
    # low-texture image: the minThFAST fallback path dominates
    rng = np.random.default_rng(5)
    low = (127 + 6 * rng.standard_normal((480, 752))).clip(0, 255).astype(np.uint8)
    kps, desc = ex(low); rk, rd = oracle.orb_extract(low)
    assert len(kps) == len(rk) and np.array_equal(desc, rd) and np.array_equal(kps["x"], rk["x"])
    # 2.4.13 blur taps behind the flag
    ex2 = ORBextractor(blur_2413=True)
    img = make_image(0)
    k2, d2 = ex2(img); rk2, rd2 = oracle.orb_extract(img, oracle.orb_cfg(blur_2413=1))
    assert np.array_equal(d2, rd2)
    ex.close(); ex2.close()


def _frames():
    a = make_image(0); b = np.roll(make_image(0), (3, 5), axis=(0, 1))  # a shifted copy: many true matches
    return a, b


def test_hamming_matrix_exact(oracle):
    rng = np.random.default_rng(0)
    A = rng.integers(0, 256, size=(517, 32), dtype=np.uint8); B = rng.integers(0, 256, size=(1003, 32), dtype=np.uint8)
    B[7] = A[11]; A[0] = 0; B[0] = 255
    D = api.hamming_matrix(A, B)
    ref = np.unpackbits(A[:, None, :] ^ B[None, :, :], axis=2).sum(axis=2)
    assert np.array_equal(D, ref.astype(np.uint16))
    assert D[0, 0] == 256 and D[11, 7] == 0
    assert oracle.descriptor_distance(A[3], B[5]) == D[3, 5]
    assert api.hamming_matrix(A[:0], B).shape == (0, 1003)


def test_search_by_bow_and_triangulation_match_indices(oracle):
    a, b = _frames()
    ex = ORBextractor()
    k1, d1 = ex(a); k2, d2 = ex(b)
    rng = np.random.default_rng(1)
    # vocabulary-node grouping stand-in: hash of the top descriptor bits gives ~100 nodes like DBoW2 at levelsup=4
    node = lambda d: (d[:, 0].astype(np.int64) * 7 + d[:, 1] // 64) % 97
    fv1, fv2 = FeatureVector(node(d1)), FeatureVector(node(d2))
    ofv1, ofv2 = oracle.FeatureVector(node(d1)), oracle.FeatureVector(node(d2))
    has1 = (rng.random(len(k1)) < 0.7).astype(np.uint8); has2 = (rng.random(len(k2)) < 0.7).astype(np.uint8)
    for nnratio, ori in [(0.7, True), (0.9, False)]:
        m = ORBmatcher(nnratio, ori)
        got, n = m.SearchByBoW_KF_Frame(d1, has1, k1["angle"], fv1, d2, k2["angle"], fv2)
        ref, rn = oracle.match_bow_kf_frame(d1, has1, k1["angle"], ofv1, d2, k2["angle"], ofv2, nnratio, ori)
        assert n == rn and np.array_equal(got, ref) and n > 20
        got, n = m.SearchByBoW_KF_KF(d1, has1, k1["angle"], fv1, d2, has2, k2["angle"], fv2)
        ref, rn = oracle.match_bow_kf_kf(d1, has1, k1["angle"], ofv1, d2, has2, k2["angle"], ofv2, nnratio, ori)
        assert n == rn and np.array_equal(got, ref)
    # triangulation: pure x-translation between the views -> F12 = [t]_x up to intrinsics
    fx, fy, cx, cy = [np.float32(v) for v in synth.EUROC_INTR]
    Kinv = np.linalg.inv(np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64))
    tx = np.array([[0, 0, 0], [0, 0, -1.0], [0, 1.0, 0]])
    F12 = (Kinv.T @ tx @ Kinv).astype(np.float32)
    sf = (1.2 ** np.arange(8)).astype(np.float32); ls2 = (sf * sf).astype(np.float32)
    v = lambda k, d, has, fv: dict(desc=d, has_mp=has, kp_xy=np.stack([k["x"], k["y"]], 1), octave=k["octave"], angle=k["angle"], fv=fv, intr=(fx, fy, cx, cy))
    for ori in (False, True):
        got = ORBmatcher(0.6, ori).SearchForTriangulation(v(k1, d1, has1, fv1), v(k2, d2, has2, fv2), F12, -5000.0, float(cy), ls2, sf)
        ref = oracle.match_triangulation(v(k1, d1, has1, ofv1), v(k2, d2, has2, ofv2), F12, -5000.0, float(cy), ls2, sf, ori)
        assert np.array_equal(got, ref)
    ex.close()


@pytest.mark.parametrize("K,fix_scale", [(60, False), (200, False), (200, True)])
def test_sim3_pose_graph_matches_oracle(oracle, K, fix_scale):
    p = synth.make_pgo(K=K, fix_scale=fix_scale)
    ref = oracle.pgo_solve(p, iterations=20)
    got = api.pgo_solve(p, iterations=20)
    # Same LM iteration count as the oracle (north_star: "after the same iteration count"), same accept / reject decision and trial count
    # in every iteration whose gain ratio is above the noise floor.  The one decision that differs in these cases (traces:
    # profiles/r2/pgo_trace.log) is iteration 4 of (K = 200, fixed scale): rho = +2.0e-12 on the oracle, -4.5e-13 on the device -- a chi2
    # difference of 2e-15 relative, eight orders below the 1e-7 noise of the reference's numeric Jacobians (central differences,
    # delta = 1e-9, G/core/base_binary_edge.hpp:147-197); the oracle accepts a step that changes nothing, the device tries ten lambdas,
    # rejects them all, and both stop after the same five iterations with the same estimate.
    assert got["iters_done"] == ref["iters_done"] >= 1 and abs(got["chi2_initial"] - ref["chi2_initial"]) <= 1e-9 * ref["chi2_initial"]
    for it in range(ref["iters_done"]):
        if abs(ref["trace"][it, 3]) > 1e-9 and abs(got["trace"][it, 3]) > 1e-9:
            assert (got["trace"][it, 3] > 0) == (ref["trace"][it, 3] > 0) and got["trace"][it, 4] == ref["trace"][it, 4], it
    assert abs(got["trace"][0, 2] - ref["trace"][0, 2]) <= 1e-4 * ref["trace"][0, 2]
    assert abs(got["chi2_final"] - ref["chi2_final"]) <= 1e-4 * ref["chi2_final"]
    scale = np.abs(ref["sim3"]).max()
    assert np.abs(got["sim3"] - ref["sim3"]).max() <= 1e-4 * scale
    assert got["chi2_final"] < 0.1 * got["chi2_initial"]
