"""The oracle's matchers (oracle/match_oracle.cpp, oracle/proj_oracle.cpp) against the REFERENCE'S OWN cslam/src/ORBmatcher.cpp, compiled
where it lies into oracle/_ref/libmatch_ref.so (oracle/Makefile `ref`).  Frame / KeyFrame / MapPoint are plain stand-ins exposing the
members the matcher source names (oracle/ref_stub/cslam/Frame.h); every search method — candidate walks, best / second-best bookkeeping,
thresholds, ratio tests, rotation histograms, mutual checks and the geometric gates in front of them — is the reference's object code.
Index-exact.  Skipped where neither /root/reference nor a prebuilt oracle/_ref is present."""
import numpy as np
import pytest

from ccm_slam_b200 import synth_match as sm

f32 = np.float32


@pytest.fixture(scope="module")
def ref(oracle):
    if oracle.ref_match() is None:
        pytest.skip("reference ORBmatcher library not available (no /root/reference, no prebuilt oracle/_ref)")
    return oracle


def _two_frames(seed, n=800):
    """descriptor sets with true matches (a few flipped bits), vocabulary-node labels and angles, as SearchByBoW sees them"""
    rng = np.random.default_rng(seed)
    d1 = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); d2 = rng.integers(0, 256, size=(n + 40, 32), dtype=np.uint8)
    perm = rng.permutation(n + 40)[: (2 * n) // 3]
    d2[perm] = sm.flip_bits(d1[: len(perm)], rng.integers(0, 60, len(perm)), rng)
    node1 = rng.integers(0, 40, n); node2 = rng.integers(0, 40, n + 40); node2[perm] = node1[: len(perm)]
    a1 = rng.uniform(0, 360, n).astype(f32); a2 = rng.uniform(0, 360, n + 40).astype(f32)
    a2[perm] = (a1[: len(perm)] + rng.normal(0, 3, len(perm)) + np.where(rng.random(len(perm)) < 0.1, 90, 0)).astype(f32) % f32(360)
    has1 = (rng.random(n) < 0.75).astype(np.uint8); has2 = (rng.random(n + 40) < 0.75).astype(np.uint8)
    return d1, d2, node1, node2, a1, a2, has1, has2


@pytest.mark.parametrize("seed,nnratio,ori", [(0, 0.7, True), (1, 0.9, False), (2, 0.6, True)])
def test_search_by_bow(ref, seed, nnratio, ori):
    d1, d2, node1, node2, a1, a2, has1, has2 = _two_frames(seed)
    fv1, fv2 = ref.FeatureVector(node1), ref.FeatureVector(node2)
    got, n = ref.match_bow_kf_frame(d1, has1, a1, fv1, d2, a2, fv2, nnratio, ori)
    want, wn = ref.ref_match_bow_kf_frame(d1, has1, a1, fv1, d2, a2, fv2, nnratio, ori)
    assert n == wn and np.array_equal(got, want) and n > 100
    got, n = ref.match_bow_kf_kf(d1, has1, a1, fv1, d2, has2, a2, fv2, nnratio, ori)
    want, wn = ref.ref_match_bow_kf_kf(d1, has1, a1, fv1, d2, has2, a2, fv2, nnratio, ori)
    assert n == wn and np.array_equal(got, want) and n > 60
    import ctypes
    a, b = np.ascontiguousarray(d1[0]), np.ascontiguousarray(d2[0])
    assert ref.ref_match().ref_descriptor_distance(a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p)) == ref.descriptor_distance(a, b)


@pytest.mark.parametrize("seed,ori", [(3, False), (4, True)])
def test_search_for_triangulation(ref, seed, ori):
    rng = np.random.default_rng(seed)
    n = 700
    d1 = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    d2 = sm.flip_bits(d1, rng.integers(0, 70, n), rng)                      # keypoint i of view 2 is keypoint i of view 1, noisily
    d2[500:] = rng.integers(0, 256, size=(n - 500, 32), dtype=np.uint8)     # ... except the last 200
    node1 = rng.integers(0, 30, n); node2 = node1.copy(); node2[650:] = rng.integers(0, 30, n - 650)
    a1 = rng.uniform(0, 360, n).astype(f32); a2 = ((a1 + rng.normal(0, 3, n) + np.where(rng.random(n) < 0.1, 90, 0)) % 360).astype(f32)
    has1 = (rng.random(n) < 0.3).astype(np.uint8); has2 = (rng.random(n) < 0.3).astype(np.uint8)   # only untracked keypoints are triangulated
    # pure x-translation between the views: F12 = K^-T [t]_x K^-1, matches lie on (nearly) the same image row
    fx, fy, cx, cy = f32(512.0), f32(512.0), f32(376.0), f32(240.0)
    xy1 = np.round(rng.uniform([20, 20], [730, 460], size=(n, 2)) * 4) / 4
    xy2 = xy1 + np.stack([rng.uniform(-40, 40, n), rng.normal(0, 1.2, n)], 1); xy2 = np.round(xy2 * 4) / 4
    oct1 = rng.integers(0, 8, n).astype(np.int32); oct2 = rng.integers(0, 8, n).astype(np.int32)
    Kinv = np.linalg.inv(np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64))
    F12 = (Kinv.T @ np.array([[0, 0, 0], [0, 0, -1.0], [0, 1.0, 0]]) @ Kinv).astype(f32)
    fv1, fv2 = ref.FeatureVector(node1), ref.FeatureVector(node2)
    v = lambda d, has, xy, oc, an, fv: dict(desc=d, has_mp=has, kp_xy=xy.astype(f32), octave=oc, angle=an, fv=fv, intr=(fx, fy, cx, cy))
    v1, v2 = v(d1, has1, xy1, oct1, a1, fv1), v(d2, has2, xy2, oct2, a2, fv2)
    # camera 1 sits at Cw in camera 2's frame; its image there is the epipole — float32, in the reference's evaluation order (:704-712)
    Cw = np.array([-8.0, 0.5, 2.0], f32)
    invz = f32(1.0) / Cw[2]
    ex = f32(f32(f32(fx * Cw[0]) * invz) + cx); ey = f32(f32(f32(fy * Cw[1]) * invz) + cy)
    sf = np.empty(8, f32); sf[0] = 1
    for i in range(1, 8):
        sf[i] = f32(sf[i - 1] * f32(1.2))
    got = ref.match_triangulation(v1, v2, F12, float(ex), float(ey), (sf * sf).astype(f32), sf, ori)
    want = ref.ref_match_triangulation(v1, v2, F12, Cw, ori)
    assert np.array_equal(got, want) and len(want) > 30


@pytest.mark.parametrize("seed,nnratio,ori", [(5, 0.9, True), (6, 0.7, False)])
def test_search_for_initialization(ref, seed, nnratio, ori):
    g2, q = sm.make_init_pair(n=900, seed=seed)
    g1 = dict(desc=q["desc"], kp_xy=q["uv"], octave=q["level"], angle=q["angle"], bounds=g2["bounds"], cols=g2["cols"], rows=g2["rows"])
    got, n = ref.search_for_initialization(g2, q, nnratio, ori)
    want, wn, prev = ref.ref_search_for_initialization(g1, g2, q["uv"], 100, nnratio, ori)
    assert n == wn and np.array_equal(got, want) and n > 100
    hit = want >= 0
    assert np.array_equal(prev[hit], g2["kp_xy"][want[hit]]) and np.array_equal(prev[~hit], q["uv"][~hit])   # vbPrevMatched refreshed (:557-560)


@pytest.mark.parametrize("seed,th,nnratio", [(7, 1.0, 0.8), (8, 3.0, 0.8), (9, 5.0, 0.6)])
def test_search_by_projection_track(ref, seed, th, nnratio):
    g = sm.make_grid(n=1200, seed=seed); q = sm.make_queries(g, m=1500, seed=seed + 50)
    rng = np.random.default_rng(seed + 100)
    m = 1500
    view_cos = np.where(rng.random(m) < 0.5, f32(0.9995), f32(0.9)).astype(f32)
    n_obs = np.where(rng.random(m) < 0.85, 3, 0).astype(np.int32); bad = (rng.random(m) < 0.05).astype(np.uint8)
    blocked = (rng.random(1200) < 0.2).astype(np.uint8)
    points = dict(desc=q["desc"], bad=bad, n_obs=n_obs, track_in_view=q["valid"], track_xy=q["uv"], track_level=q["level"], track_view_cos=view_cos)
    # the oracle takes the window the reference derives: RadiusByViewingCos [* th] * mvScaleFactors[level], in float32 (:86-93)
    sf = np.empty(8, f32); sf[0] = 1
    for i in range(1, 8):
        sf[i] = f32(sf[i - 1] * f32(1.2))
    r = np.where(view_cos > f32(0.998), f32(2.5), f32(4.0)).astype(f32)
    if th != 1.0:
        r = (r * f32(th)).astype(f32)
    oq = dict(q, valid=(q["valid"].astype(bool) & ~bad.astype(bool)).astype(np.uint8), radius=(r * sf[q["level"]]).astype(f32))
    got, n = ref.search_by_projection_track(g, oq, (n_obs > 0).astype(np.uint8), blocked, nnratio)
    want, wn = ref.ref_search_by_projection_track(g, points, blocked, th, nnratio)
    assert n == wn and np.array_equal(got, want) and n > 200
