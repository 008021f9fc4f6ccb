"""The oracle's ORB extractor (oracle/orb_oracle.cpp) against the REFERENCE'S OWN cslam/src/ORBextractor.cpp, compiled where it lies
into oracle/_ref/orb_ref_cli (oracle/Makefile `ref`) against the stand-in headers of oracle/ref_stub/.  The five OpenCV primitives the
reference calls (FAST, resize, GaussianBlur, copyMakeBorder, fastAtan2) are supplied from the oracle's restatements, which
tests/test_oracle_orb.py pins to cv2 4.13 — so this test holds the oracle's EXTRACTOR LOGIC (scale tables, umax, pyramid, 30-px cells with
the 20 -> 7 fallback, quadtree distribution, IC_Angle, rotated BRIEF sampling with the float cos / sin, scaling, output order) to the
reference's object code on identical primitives: keypoints and descriptors bit for bit.

One thing the reference does not define: DistributeOctTree orders equal-size nodes by heap address (ORBextractor.cpp:852).  With a
monotone allocator (addresses grow with allocation order) the reference equals the oracle exactly; on glibc's allocator the reference
differs from ITSELF-under-bump in a handful of keypoints per frame — measured below, so the claim "bit-exact" is stated against the
monotone-allocator behaviour.  Skipped where neither /root/reference nor a prebuilt oracle/_ref is present."""
import numpy as np
import pytest

from ccm_slam_b200.synth_images import make_image

FIELDS = ("x", "y", "size", "angle", "response", "octave")


@pytest.fixture(scope="module")
def ref(oracle):
    if oracle.ref_orb_cli() is None:
        pytest.skip("reference ORBextractor program not available (no /root/reference, no prebuilt oracle/_ref)")
    return oracle


def _same(a, b):
    (ka, da), (kb, db) = a, b
    return len(ka) == len(kb) and all(np.array_equal(ka[f], kb[f]) for f in FIELDS) and np.array_equal(da, db)


@pytest.mark.parametrize("seed,w,h", [(0, 752, 480), (1, 752, 480), (2, 640, 480), (3, 376, 240), (10, 752, 480), (11, 1024, 768)])
def test_extractor_equals_the_reference_code(ref, seed, w, h):
    img = make_image(seed, w, h)
    got, want = ref.orb_extract(img), ref.ref_orb_extract(img)
    assert len(want[0]) > 500 and _same(got, want)


@pytest.mark.parametrize("cfg", [dict(nfeatures=500), dict(nfeatures=2000), dict(nlevels=4), dict(scale_factor=1.5, nlevels=5),
                                 dict(ini_th=40, min_th=12), dict(blur_2413=1)])
def test_other_configurations(ref, cfg):
    img = make_image(4, 752, 480)
    c = ref.orb_cfg(**cfg)
    got, want = ref.orb_extract(img, c), ref.ref_orb_extract(img, c)
    assert len(want[0]) > 100 and _same(got, want)


def test_degenerate_images(ref):
    flat = np.full((480, 752), 127, np.uint8)
    got, want = ref.orb_extract(flat), ref.ref_orb_extract(flat)
    assert len(got[0]) == len(want[0]) == 0
    rng = np.random.default_rng(5)
    low = (127 + 6 * rng.standard_normal((480, 752))).clip(0, 255).astype(np.uint8)     # the minThFAST fallback dominates
    assert _same(ref.orb_extract(low), ref.ref_orb_extract(low))
    noise = rng.integers(0, 256, size=(240, 376), dtype=np.uint8)                        # far more corners than the quota
    assert _same(ref.orb_extract(noise), ref.ref_orb_extract(noise))


def test_the_reference_itself_depends_on_the_allocator(ref):
    """documented, not desired: on glibc's allocator the reference's quadtree splits equal-size nodes in heap-address order"""
    img = make_image(0, 752, 480)
    kb, db = ref.ref_orb_extract(img, allocator="bump")
    km, dm = ref.ref_orb_extract(img, allocator="malloc")
    key = lambda k: set(zip(k["octave"].tolist(), k["x"].tolist(), k["y"].tolist()))
    common = key(kb) & key(km)
    assert len(common) >= 0.97 * len(kb) and abs(len(kb) - len(km)) <= 0.01 * len(kb)
    # a keypoint both runs keep is the same keypoint: angle, response, size and descriptor agree
    ib = {(int(o), float(x), float(y)): i for i, (o, x, y) in enumerate(zip(kb["octave"], kb["x"], kb["y"]))}
    im = {(int(o), float(x), float(y)): i for i, (o, x, y) in enumerate(zip(km["octave"], km["x"], km["y"]))}
    for k in list(common)[:400]:
        a, b = ib[k], im[k]
        assert kb["angle"][a] == km["angle"][b] and kb["response"][a] == km["response"][b] and np.array_equal(db[a], dm[b])
