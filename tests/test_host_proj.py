"""CPU suite, host logic of the product for SURVEY.md §8(f) ranks 2-3: the selection half of every projection-guided matcher
(ccm_select_*, fed a numpy distance matrix), GetFeaturesInArea and the BowVector / FeatureVector assembly, against the oracle.
No device work is involved (the device half — k_hamming, k_voc_descend — is covered by tests/test_gpu_widen.py)."""
import numpy as np
import pytest

from ccm_slam_b200 import api
from ccm_slam_b200 import synth_match as sm
from ccm_slam_b200.frontend import GetFeaturesInArea, ORBmatcher, bow_assemble


def dist(q, g):
    return np.unpackbits(np.asarray(q["desc"])[:, None, :] ^ np.asarray(g["desc"])[None, :, :], axis=2).sum(axis=2).astype(np.uint16)


@pytest.fixture(scope="module", params=[(0, 1000, 1500, 3.0), (1, 2000, 3000, 7.0), (2, 300, 200, 15.0), (3, 1200, 1800, 6.0)])
def case(request):
    seed, n, m, th = request.param
    g = sm.make_grid(n=n, seed=10 + seed, clustered=seed != 2)
    q = sm.make_queries(g, m=m, seed=20 + seed, th=th)
    if seed == 3:                                   # every window full of candidates at identical distances: the visiting order decides
        g, q = sm.tie_storm(g, q, pool=6, seed=40)
    rng = np.random.default_rng(30 + seed)
    return dict(g=g, q=q, D=dist(q, g), has_obs=(rng.random(m) < 0.85).astype(np.uint8), blocked=(rng.random(n) < 0.2).astype(np.uint8),
                existing=np.where(rng.random(m) < 0.15, rng.integers(0, n, m), -1).astype(np.int32))


def test_features_in_area(oracle):
    g = sm.make_grid(n=1500, seed=1)
    rng = np.random.default_rng(2)
    total = 0
    for k in range(400):
        j = int(rng.integers(0, 1500))
        x, y = g["kp_xy"][j] + rng.normal(0, 4, 2).astype(np.float32)
        r = float(rng.choice([0.25, 2.0, 7.5, 30.0, 400.0, 5000.0]))
        lo, hi = (int(rng.integers(-1, 7)), int(rng.integers(-1, 8))) if k % 2 else (-1, -1)
        got = GetFeaturesInArea(g, x, y, r, lo, hi); ref = oracle.features_in_area(g, x, y, r, lo, hi)
        assert np.array_equal(got, ref)
        total += len(ref)
    assert total > 10000
    x0, y0, x1, y1 = g["bounds"]
    for (x, y) in [(x0 - 50, 100), (x1 + 50, 100), (100, y0 - 50), (100, y1 + 50), (x0 - 5, y0 - 5), (x1 + 1, y1 + 1)]:
        assert np.array_equal(GetFeaturesInArea(g, x, y, 20.0), oracle.features_in_area(g, x, y, 20.0))
    empty = dict(g, desc=g["desc"][:0], kp_xy=g["kp_xy"][:0], octave=g["octave"][:0], angle=g["angle"][:0])
    assert len(GetFeaturesInArea(empty, 10.0, 10.0, 50.0)) == 0


@pytest.mark.parametrize("nnratio", [0.8, 0.6])
def test_select_track(oracle, case, nnratio):
    got, n = ORBmatcher(nnratio).SearchByProjection_Track(case["g"], case["q"], case["has_obs"], case["blocked"], D=case["D"])
    ref, rn = oracle.search_by_projection_track(case["g"], case["q"], case["has_obs"], case["blocked"], nnratio)
    assert n == rn and np.array_equal(got, ref) and n > 20


@pytest.mark.parametrize("reloc,orb_dist,ori", [(False, 100, True), (False, 100, False), (True, 64, True), (True, 100, False)])
def test_select_frame(oracle, case, reloc, orb_dist, ori):
    got, n = ORBmatcher(0.9, ori).SearchByProjection_Frame(case["g"], case["q"], case["has_obs"], case["blocked"], reloc, orb_dist, D=case["D"])
    ref, rn = oracle.search_by_projection_frame(case["g"], case["q"], case["has_obs"], case["blocked"], reloc, orb_dist, ori)
    assert n == rn and np.array_equal(got, ref) and n > 20


def test_select_sim3_and_fuse(oracle, case):
    m = ORBmatcher()
    best, mof, n = m.SearchByProjection_Sim3(case["g"], case["q"], case["blocked"], case["existing"], D=case["D"])
    rbest, rmof, rn = oracle.search_by_projection_sim3(case["g"], case["q"], case["blocked"], case["existing"])
    assert n == rn and np.array_equal(best, rbest) and np.array_equal(mof, rmof) and n > 20
    for w in (None, sm.INV_LEVEL_SIGMA2):
        best, n = m.Fuse(case["g"], case["q"], w, D=case["D"])
        rbest, rn = oracle.fuse_search(case["g"], case["q"], w)
        assert n == rn and np.array_equal(best, rbest) and n > 20


def test_select_by_sim3(oracle):
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
    got, n = ORBmatcher().SearchBySim3(g1, g2, q12, q21, D12=dist(q12, g2), D21=dist(q21, g1))
    ref, rn = oracle.search_by_sim3(g1, g2, q12, q21)
    assert n == rn and np.array_equal(got, ref) and n > 150


@pytest.mark.parametrize("nnratio,ori", [(0.9, True), (0.7, False)])
def test_select_for_initialization(oracle, nnratio, ori):
    g2, q = sm.make_init_pair(n=1500, seed=5)
    got, n = ORBmatcher(nnratio, ori).SearchForInitialization(g2, q, D=dist(q, g2))
    ref, rn = oracle.search_for_initialization(g2, q, nnratio, ori)
    assert n == rn and np.array_equal(got, ref) and n > 150


def test_select_edge_cases(oracle):
    g = sm.make_grid(n=200, seed=3); q = sm.make_queries(g, m=50, seed=4)
    m = ORBmatcher(0.8, True)
    # no queries / no valid queries / no features
    q0 = {k: v[:0] for k, v in q.items()}
    got, n = m.SearchByProjection_Track(g, q0, np.zeros(0, np.uint8), np.zeros(200, np.uint8), D=np.zeros((0, 200), np.uint16))
    assert n == 0 and (got == -1).all()
    qi = dict(q, valid=np.zeros(50, np.uint8))
    best, n = m.Fuse(g, qi, None, D=dist(qi, g))
    assert n == 0 and (best == -1).all()
    g0 = dict(g, desc=g["desc"][:0], kp_xy=g["kp_xy"][:0], octave=g["octave"][:0], angle=g["angle"][:0])
    got, n = m.SearchByProjection_Frame(g0, q, np.ones(50, np.uint8), np.zeros(0, np.uint8), D=np.zeros((50, 0), np.uint16))
    assert n == 0 and len(got) == 0
    # every feature shut: nothing can match
    got, n = m.SearchByProjection_Frame(g, q, np.ones(50, np.uint8), np.ones(200, np.uint8), D=dist(q, g))
    assert n == 0
    # malformed input is refused, not read
    with pytest.raises(api.CCMError):
        m.Fuse(dict(g, cols=0), q, None, D=dist(q, g))


@pytest.mark.parametrize("scoring,weighting,levelsup", [(0, 0, 4), (0, 0, 1), (1, 1, 2), (5, 0, 1), (5, 2, 1), (2, 3, 0)])
def test_bow_assemble(oracle, scoring, weighting, levelsup):
    voc = sm.make_vocabulary(k=10, L=3, seed=5, scoring=scoring, weighting=weighting)
    feat = sm.make_voc_features(voc, n=1500, seed=6)
    V = oracle.Vocabulary(voc)
    ref = V.transform(feat, levelsup)
    got = bow_assemble(scoring, weighting, ref["word"], ref["weight"], ref["node"])
    for k in ("bow_id", "bow_val", "fv_node_id", "fv_node_ptr", "fv_feat"):
        assert np.array_equal(got[k], ref[k]), k        # doubles bit-exact: same summation order
    assert len(got["bow_id"]) > 100
    V.close()
    e = bow_assemble(scoring, weighting, np.zeros(0, np.uint32), np.zeros(0), np.zeros(0, np.uint32))
    assert len(e["bow_id"]) == 0 and list(e["fv_node_ptr"]) == [0]
