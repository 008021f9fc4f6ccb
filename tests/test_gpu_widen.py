"""GPU parity tests of the rows SURVEY.md §8(f) marks "next" (ranks 2-3), through the C ABI against the CPU oracle:
the projection-guided matchers (SearchByProjection x4, Fuse x2, SearchBySim3) index-exact, the DBoW2 transform word / node
exact and BowVector values bit-exact."""
import numpy as np
import pytest

from ccm_slam_b200 import api
from ccm_slam_b200 import synth_match as sm
from ccm_slam_b200.frontend import ORBextractor, ORBmatcher, ORBVocabulary
from ccm_slam_b200.synth_images import make_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _dev():
    assert api.device_count() > 0
    api.init(0)


@pytest.fixture(scope="module", params=[(0, 1000, 1500, 3.0), (1, 2000, 4000, 7.0), (2, 300, 33, 15.0), (3, 1200, 1800, 6.0)])
def case(request):
    seed, n, m, th = request.param
    g = sm.make_grid(n=n, seed=40 + seed, clustered=seed != 2)
    q = sm.make_queries(g, m=m, seed=50 + seed, th=th)
    if seed == 3:       # ties in every window (six nearby descriptor patterns): the visiting order alone decides — the case that matters
        g, q = sm.tie_storm(g, q, pool=6, seed=70)      # most for an on-device selection (CCM_MATCH_WINDOW=1)
    rng = np.random.default_rng(60 + seed)
    return dict(g=g, q=q, has_obs=(rng.random(m) < 0.85).astype(np.uint8), blocked=(rng.random(n) < 0.2).astype(np.uint8),
                existing=np.where(rng.random(m) < 0.15, rng.integers(0, n, m), -1).astype(np.int32))


def test_search_by_projection_all_overloads(oracle, case):
    g, q = case["g"], case["q"]
    l0 = api.kernel_launches()
    for nnratio in (0.8, 0.6):
        got, n = ORBmatcher(nnratio).SearchByProjection_Track(g, q, case["has_obs"], case["blocked"])
        ref, rn = oracle.search_by_projection_track(g, q, case["has_obs"], case["blocked"], nnratio)
        assert n == rn and np.array_equal(got, ref) and n > 5
    for reloc, orb_dist, ori in [(False, 100, True), (False, 100, False), (True, 64, True), (True, 100, False)]:
        got, n = ORBmatcher(0.9, ori).SearchByProjection_Frame(g, q, case["has_obs"], case["blocked"], reloc, orb_dist)
        ref, rn = oracle.search_by_projection_frame(g, q, case["has_obs"], case["blocked"], reloc, orb_dist, ori)
        assert n == rn and np.array_equal(got, ref) and n > 5
    best, mof, n = ORBmatcher().SearchByProjection_Sim3(g, q, case["blocked"], case["existing"])
    rbest, rmof, rn = oracle.search_by_projection_sim3(g, q, case["blocked"], case["existing"])
    assert n == rn and np.array_equal(best, rbest) and np.array_equal(mof, rmof)
    for w in (None, sm.INV_LEVEL_SIGMA2):
        best, n = ORBmatcher().Fuse(g, q, w)
        rbest, rn = oracle.fuse_search(g, q, w)
        assert n == rn and np.array_equal(best, rbest) and n > 5
    assert api.kernel_launches() >= l0 + 9     # one distance-matrix launch per call: the device did the distances


def test_search_by_sim3_mutual(oracle):
    rng = np.random.default_rng(9)
    g1 = sm.make_grid(n=900, seed=7); g2 = sm.make_grid(n=950, seed=8)
    share = rng.permutation(900)[:500]
    g2["desc"][:500] = sm.flip_bits(g1["desc"][share], rng.integers(0, 30, 500), rng)
    g2["kp_xy"][:500] = g1["kp_xy"][share] + rng.normal(0, 1.5, (500, 2)).astype(np.float32)
    g2["octave"][:500] = g1["octave"][share]

    def queries(src_g, dst_g, ps, pd):
        m = src_g["desc"].shape[0]
        uv = rng.uniform(0, 700, (m, 2)).astype(np.float32); level = src_g["octave"].copy()
        uv[ps] = dst_g["kp_xy"][pd] + rng.normal(0, 1.0, (len(ps), 2)).astype(np.float32)
        return dict(valid=(rng.random(m) < 0.8).astype(np.uint8), uv=uv, radius=(np.float32(7.5) * sm.SCALE_FACTORS[level]).astype(np.float32),
                    level=level, desc=src_g["desc"])
    q12 = queries(g1, g2, share, np.arange(500)); q21 = queries(g2, g1, np.arange(500), share)
    got, n = ORBmatcher().SearchBySim3(g1, g2, q12, q21)
    ref, rn = oracle.search_by_sim3(g1, g2, q12, q21)
    assert n == rn and np.array_equal(got, ref) and n > 150


def test_search_for_initialization(oracle):
    g2, q = sm.make_init_pair(n=1500, seed=5)
    for nnratio, ori in [(0.9, True), (0.7, False)]:
        got, n = ORBmatcher(nnratio, ori).SearchForInitialization(g2, q)
        ref, rn = oracle.search_for_initialization(g2, q, nnratio, ori)
        assert n == rn and np.array_equal(got, ref) and n > 150


def test_projection_matchers_on_extracted_frames(oracle):
    """End to end on real extractor output: frame b is frame a shifted by (5, 3) px; every keypoint of a is 'projected' to its
    shifted position and searched for in b's grid."""
    a = make_image(0); b = np.roll(a, (3, 5), axis=(0, 1))
    ex = ORBextractor()
    k1, d1 = ex(a); k2, d2 = ex(b)
    ex.close()
    g = dict(desc=d2, kp_xy=np.stack([k2["x"], k2["y"]], 1), octave=k2["octave"], angle=k2["angle"], bounds=(0.0, 0.0, 752.0, 480.0), cols=75, rows=48)
    lv = k1["octave"].astype(np.int32)
    q = dict(valid=np.ones(len(k1), np.uint8), uv=np.stack([k1["x"] + 5, k1["y"] + 3], 1).astype(np.float32),
             radius=(np.float32(7.0) * sm.SCALE_FACTORS[lv]).astype(np.float32), level=lv, desc=d1, angle=k1["angle"])
    ones = np.ones(len(k1), np.uint8); none = np.zeros(len(k2), np.uint8)
    got, n = ORBmatcher(0.9, True).SearchByProjection_Frame(g, q, ones, none)
    ref, rn = oracle.search_by_projection_frame(g, q, ones, none, False, 100, True)
    assert n == rn and np.array_equal(got, ref) and n > 300
    got, n = ORBmatcher(0.8).SearchByProjection_Track(g, q, ones, none)
    ref, rn = oracle.search_by_projection_track(g, q, ones, none, 0.8)
    assert n == rn and np.array_equal(got, ref) and n > 300


@pytest.mark.parametrize("k,L,scoring,weighting,levelsup,n", [(10, 3, 0, 0, 1, 2000), (10, 4, 0, 0, 2, 1000), (6, 3, 1, 1, 2, 777), (4, 5, 5, 2, 4, 1),
                                                             (20, 2, 0, 3, 0, 300), (3, 6, 0, 0, 4, 1000)])
def test_voc_transform(oracle, k, L, scoring, weighting, levelsup, n):
    voc = sm.make_vocabulary(k=k, L=L, seed=70 + k, scoring=scoring, weighting=weighting)
    feat = sm.make_voc_features(voc, n=n, seed=80 + L)
    V = ORBVocabulary(voc); R = oracle.Vocabulary(voc)
    l0 = api.kernel_launches()
    got = V.transform(feat, levelsup); ref = R.transform(feat, levelsup)
    assert api.kernel_launches() == l0 + 1
    for key in ("word", "node", "weight", "bow_id", "bow_val", "fv_node_id", "fv_node_ptr", "fv_feat"):
        assert np.array_equal(got[key], ref[key]), key
    # a second batch through the same handle, and the empty batch
    feat2 = sm.make_voc_features(voc, n=max(1, n // 3), seed=99)
    g2 = V.transform(feat2, levelsup); r2 = R.transform(feat2, levelsup)
    assert np.array_equal(g2["word"], r2["word"]) and np.array_equal(g2["bow_val"], r2["bow_val"])
    e = V.transform(feat[:0], levelsup)
    assert len(e["bow_id"]) == 0 and list(e["fv_node_ptr"]) == [0]
    assert V.words() == int(np.asarray(voc["is_leaf"]).sum())
    V.close(); R.close()


def test_voc_feeds_search_by_bow(oracle):
    """extractor -> vocabulary -> SearchByBoW: the FeatureVector the device tree produces drives the matcher exactly as the oracle's"""
    from ccm_slam_b200.frontend import FeatureVector

    def fv_of(cls, t):   # the transform's own FeatureVector (stopped words left out), in the flattened form the matchers take
        f = cls.__new__(cls)
        f.node_id, f.node_ptr, f.feat = t["fv_node_id"].astype(np.uint32), t["fv_node_ptr"].astype(np.int32), t["fv_feat"].astype(np.uint32)
        return f
    a = make_image(0); b = np.roll(a, (3, 5), axis=(0, 1))
    ex = ORBextractor()
    k1, d1 = ex(a); k2, d2 = ex(b)
    ex.close()
    voc = sm.make_vocabulary(k=10, L=3, seed=1)
    V = ORBVocabulary(voc); R = oracle.Vocabulary(voc)
    t1, t2 = V.transform(d1, 2), V.transform(d2, 2)
    r1, r2 = R.transform(d1, 2), R.transform(d2, 2)
    assert np.array_equal(t1["node"], r1["node"]) and np.array_equal(t2["node"], r2["node"])
    has = np.ones(len(k1), np.uint8)
    got, n = ORBmatcher(0.7, True).SearchByBoW_KF_Frame(d1, has, k1["angle"], fv_of(FeatureVector, t1), d2, k2["angle"], fv_of(FeatureVector, t2))
    ref, rn = oracle.match_bow_kf_frame(d1, has, k1["angle"], fv_of(oracle.FeatureVector, r1), d2, k2["angle"], fv_of(oracle.FeatureVector, r2), 0.7, True)
    assert n == rn and np.array_equal(got, ref) and n > 100
    V.close(); R.close()


def test_gpu_matches_proj_and_voc_fixtures():
    """the committed golden fixtures (tests/golden/proj_matchers.npz, voc_k6_L3.npz) through the C ABI, device distances / descent"""
    from tests.test_golden import check_proj_fixture, load, unpack_voc

    class M:
        track = staticmethod(lambda G, Q, ho, bl, nn: ORBmatcher(nn).SearchByProjection_Track(G, Q, ho, bl))
        frame = staticmethod(lambda G, Q, ho, bl, reloc, od, ori: ORBmatcher(0.9, ori).SearchByProjection_Frame(G, Q, ho, bl, reloc, od))
        sim3proj = staticmethod(lambda G, Q, fm, ex: ORBmatcher().SearchByProjection_Sim3(G, Q, fm, ex))
        fuse = staticmethod(lambda G, Q, w: ORBmatcher().Fuse(G, Q, w))
        mutual = staticmethod(lambda G1, G2, Q12, Q21: ORBmatcher().SearchBySim3(G1, G2, Q12, Q21))
    check_proj_fixture(load("proj_matchers.npz"), M)
    g = load("voc_k6_L3.npz")
    V = ORBVocabulary(unpack_voc(g))
    for levelsup in (1, 2):
        r = V.transform(g["feat"], levelsup)
        for k, v in r.items():
            assert np.array_equal(v, g["l%d_%s" % (levelsup, k)]), k
    V.close()
