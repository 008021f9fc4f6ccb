"""CPU suite: the oracle's projection-guided matchers and DBoW2 transform against the
independent pure-Python witnesses of tests/witness_match.py.  Index-exact; BowVector values bit-exact (same summation order)."""
import numpy as np
import pytest

from ccm_slam_b200 import synth_match as sm
from tests import witness_match as wm


@pytest.fixture(scope="module")
def case():
    g = sm.make_grid(n=700, seed=3)
    q = sm.make_queries(g, m=900, seed=4, th=4.0)
    rng = np.random.default_rng(5)
    return dict(g=g, q=q, has_obs=(rng.random(900) < 0.85).astype(np.uint8), blocked=(rng.random(700) < 0.2).astype(np.uint8),
                existing=np.where(rng.random(900) < 0.15, rng.integers(0, 700, 900), -1).astype(np.int32))


def test_features_in_area_order_and_edges(oracle):
    g = sm.make_grid(n=500, seed=1)
    G = wm.GridW(g)
    rng = np.random.default_rng(2)
    total = 0
    for k in range(200):
        j = int(rng.integers(0, 500))
        x, y = g["kp_xy"][j] + rng.normal(0, 3, 2).astype(np.float32)
        r = np.float32(rng.choice([0.25, 2.0, 7.5, 30.0, 400.0]))
        if k % 3 == 0:
            lo, hi = int(rng.integers(-1, 7)), int(rng.integers(0, 8))
            ref = G.in_area(x, y, r, lo, hi) if hi >= 0 else G.in_area(x, y, r)
            got = oracle.features_in_area(g, x, y, r, lo, hi)
        else:
            ref = G.in_area(x, y, r); got = oracle.features_in_area(g, x, y, r)
        assert list(got) == ref
        total += len(ref)
    assert total > 2000
    # windows that leave the image on each side, and a window exactly touching a keypoint (strict '<')
    x0, y0, x1, y1 = g["bounds"]
    for (x, y) in [(x0 - 50, 100), (x1 + 50, 100), (100, y0 - 50), (100, y1 + 50), (x0 - 5, y0 - 5), (x1 + 1, y1 + 1)]:
        assert list(oracle.features_in_area(g, x, y, 20.0)) == G.in_area(x, y, 20.0)
    j = int(np.flatnonzero(G.ingrid)[0]); px, py = g["kp_xy"][j]
    assert j not in oracle.features_in_area(g, px + 2.0, py, 2.0) and j in oracle.features_in_area(g, px + 2.0, py, 2.25)
    # keypoints outside the bounds never come back, whatever the window
    everything = set(oracle.features_in_area(g, 300.0, 200.0, 5000.0).tolist())
    assert everything == set(np.flatnonzero(G.ingrid).tolist()) and len(everything) < 500


@pytest.mark.parametrize("nnratio", [0.8, 0.6])
def test_search_by_projection_track(oracle, case, nnratio):
    got, n = oracle.search_by_projection_track(case["g"], case["q"], case["has_obs"], case["blocked"], nnratio)
    ref, rn = wm.search_track(case["g"], case["q"], case["has_obs"], case["blocked"], nnratio)
    assert n == rn and np.array_equal(got, ref) and n > 200
    # shielding matters: with no query shielding its feature more assignments happen (later queries overwrite)
    _, n0 = oracle.search_by_projection_track(case["g"], case["q"], np.zeros(900, np.uint8), case["blocked"], nnratio)
    assert n0 >= n


@pytest.mark.parametrize("reloc,orb_dist,ori", [(False, 100, True), (False, 100, False), (True, 64, True), (True, 100, False)])
def test_search_by_projection_frame(oracle, case, reloc, orb_dist, ori):
    got, n = oracle.search_by_projection_frame(case["g"], case["q"], case["has_obs"], case["blocked"], reloc, orb_dist, ori)
    ref, rn = wm.search_frame(case["g"], case["q"], case["has_obs"], case["blocked"], reloc, orb_dist, ori)
    assert n == rn and np.array_equal(got, ref) and n > 150
    if ori:
        assert (got == -2).sum() > 0   # the orientation check removed something


def test_search_by_projection_sim3_and_fuse(oracle, case):
    best, mof, n = oracle.search_by_projection_sim3(case["g"], case["q"], case["blocked"], case["existing"])
    rbest, rmof, rn = wm.search_sim3proj(case["g"], case["q"], case["blocked"], case["existing"])
    assert n == rn and np.array_equal(best, rbest) and np.array_equal(mof, rmof) and n > 100
    assert ((best >= 0) & (case["existing"] >= 0)).sum() > 5      # the remap branch is exercised
    for w in (None, sm.INV_LEVEL_SIGMA2):
        best, n = oracle.fuse_search(case["g"], case["q"], w)
        rbest, rn = wm.fuse_search(case["g"], case["q"], w)
        assert n == rn and np.array_equal(best, rbest) and n > 100
    # the chi-square gate of Fuse(kf, points) bites
    assert oracle.fuse_search(case["g"], case["q"], sm.INV_LEVEL_SIGMA2)[1] < oracle.fuse_search(case["g"], case["q"], None)[1]


def test_search_by_sim3_mutual(oracle):
    g1 = sm.make_grid(n=400, seed=7); g2 = sm.make_grid(n=420, seed=8)
    # keyframe 2 sees perturbed copies of half of keyframe 1's features
    rng = np.random.default_rng(9)
    share = rng.permutation(400)[:200]
    g2["desc"][:200] = sm.flip_bits(g1["desc"][share], rng.integers(0, 30, 200), rng)
    g2["kp_xy"][:200] = g1["kp_xy"][share] + rng.normal(0, 1.5, (200, 2)).astype(np.float32)
    g2["octave"][:200] = g1["octave"][share]

    def queries(src_g, dst_g, pairs_src, pairs_dst):   # one query per source feature
        m = src_g["desc"].shape[0]
        uv = rng.uniform(0, 700, (m, 2)).astype(np.float32); level = src_g["octave"].copy()
        uv[pairs_src] = dst_g["kp_xy"][pairs_dst] + rng.normal(0, 1.0, (len(pairs_src), 2)).astype(np.float32)
        valid = (rng.random(m) < 0.8).astype(np.uint8)
        return dict(valid=valid, uv=uv, radius=(np.float32(7.5) * sm.SCALE_FACTORS[level]).astype(np.float32), level=level, desc=src_g["desc"])
    q12 = queries(g1, g2, share, np.arange(200)); q21 = queries(g2, g1, np.arange(200), share)
    got, n = oracle.search_by_sim3(g1, g2, q12, q21)
    ref, rn = wm.search_by_sim3(g1, g2, q12, q21)
    assert n == rn and np.array_equal(got, ref) and n > 60


@pytest.mark.parametrize("nnratio,ori", [(0.9, True), (0.7, False)])
def test_search_for_initialization(oracle, nnratio, ori):
    g2, q = sm.make_init_pair(n=500, seed=3)
    got, n = oracle.search_for_initialization(g2, q, nnratio, ori)
    ref, rn = wm.search_init(g2, q, nnratio, ori)
    assert n == rn and np.array_equal(got, ref) and n > 60
    assert (got[q["level"] > 0] == -1).all()                    # only octave-0 keypoints of F1 are searched
    hit = got[got >= 0]
    assert len(np.unique(hit)) == len(hit)                      # a keypoint of F2 stays with one keypoint of F1
    assert (g2["octave"][hit] == 0).all()


@pytest.mark.parametrize("scoring,weighting,levelsup", [(0, 0, 1), (0, 0, 4), (1, 1, 2), (5, 0, 1), (2, 2, 0), (0, 3, 1)])
def test_voc_transform(oracle, scoring, weighting, levelsup):
    voc = sm.make_vocabulary(k=6, L=3, seed=11, scoring=scoring, weighting=weighting)
    feat = sm.make_voc_features(voc, n=400, seed=12)
    V = oracle.Vocabulary(voc)
    got = V.transform(feat, levelsup)
    ref = wm.voc_transform(voc, feat, levelsup)
    assert [(int(a), int(b), float(c)) for a, b, c in zip(got["word"], got["node"], got["weight"])] == ref["per"]
    assert list(got["bow_id"]) == ref["bow_id"] and list(got["bow_val"]) == ref["bow_val"]     # bit-exact doubles
    assert list(got["fv_node_id"]) == list(ref["fv"].keys())
    for a, k in enumerate(ref["fv"]):
        assert list(got["fv_feat"][got["fv_node_ptr"][a]:got["fv_node_ptr"][a + 1]]) == ref["fv"][k]
    if scoring == 0:
        assert abs(got["bow_val"].sum() - 1.0) < 1e-12
    assert len(got["bow_id"]) > 50 and (got["weight"] == 0).sum() > 0   # stopped words occur and are skipped
    if levelsup >= 3:
        assert list(got["fv_node_id"]) == [0]                            # L - levelsup <= 0: everything hangs off the root
    V.close()


def test_voc_known_answers(oracle):
    # two levels, k = 2: children of the root are all-zero and all-one; their children differ in one byte
    z, o = np.zeros(32, np.uint8), np.full(32, 255, np.uint8)
    def with_byte(b, v):
        d = b.copy(); d[0] = v; return d
    desc = np.stack([z, z, o, with_byte(z, 0x0F), with_byte(z, 0xF0), with_byte(o, 0x0F), with_byte(o, 0x0F)])
    voc = dict(k=2, L=2, scoring=0, weighting=0, parent=[0, 0, 0, 1, 1, 2, 2], is_leaf=[0, 0, 0, 1, 1, 1, 1], desc=desc,
               weight=[0, 0, 0, 1.0, 2.0, 4.0, 8.0])
    V = oracle.Vocabulary(voc)
    f = np.stack([with_byte(z, 0x0F), with_byte(z, 0xF1), o, with_byte(z, 0x3C)])
    r = V.transform(f, 1)
    assert list(r["word"]) == [0, 1, 2, 0]          # last one: equidistant to both leaves -> the first wins; 'o': duplicate siblings -> first
    assert list(r["node"]) == [1, 1, 2, 1]          # level L - levelsup = 1
    assert list(r["bow_id"]) == [0, 1, 2] and np.allclose(r["bow_val"], np.array([2.0, 2.0, 4.0]) / 8.0)
    assert list(r["fv_node_id"]) == [1, 2] and list(r["fv_feat"]) == [0, 1, 3, 2]
    V.close()
