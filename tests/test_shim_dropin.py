"""CPU suite: the reference-side shims as a drop-in for cslam::ORBmatcher.

shim/ORBmatcher_shim.cpp + shim/ORBmatcher_proj_shim.cpp are compiled against the reference's own cslam/ORBmatcher.h (so every
member signature is checked by the compiler) and the stand-in Frame / KeyFrame / MapPoint of oracle/ref_stub, behind the same C wrappers
as the reference's ORBmatcher.cpp (oracle/Makefile: _ref/libmatch_shim.so next to _ref/libmatch_ref.so).  Every scene of
tests/test_oracle_vs_reference_matchers.py is then pushed through BOTH implementations of the class — same objects, same calls — and
the results must be identical: match arrays, counts, the RemapMapPointMatch call list, the refreshed vbPrevMatched, what Fuse adds and
replaces.  That exercises the shims' own code: pose algebra, projection, image / distance / viewing-angle gates, PredictScale, window
radii, the write-back epilogues.  Without a GPU the device half of the six ccm_search_* entry points is a link-time double (CPU Hamming
matrix + the library's own ccm_select_*, oracle/ccm_search_double.cpp), so all eleven search methods and DescriptorDistance run here.  Skipped where the reference tree or the product library is absent."""
import ctypes as C

import numpy as np
import pytest

from tests import test_oracle_vs_reference_matchers as T

DEVICE_ONLY = set()     # every method of the class has a host half behind the double since ccm_select_bow_* / ccm_select_triangulation


def same(a, b):
    if isinstance(a, (tuple, list)):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            same(x, y)
    elif isinstance(a, np.ndarray):
        assert np.array_equal(a, b)
    else:
        assert a == b


class SideBySide:
    """the oracle module, with every call into the matcher class made through both libraries"""

    def __init__(self, oracle):
        self._o = oracle
        self.calls = 0

    def __getattr__(self, name):
        f = getattr(self._o, name)
        if not name.startswith("ref_") or name in DEVICE_ONLY or name == "ref_match" or not callable(f):
            return f

        def both(*a, **k):
            r = f(*a, **k)
            with self._o.matcher_side("shim"):
                s = f(*a, **k)
            same(r, s)
            self.calls += 1
            return r
        return both


@pytest.fixture(scope="module")
def side(oracle):
    if oracle.ref_match() is None:
        pytest.skip("reference tree absent and no prebuilt oracle/_ref/libmatch_ref.so")
    with oracle.matcher_side("shim"):
        if oracle.ref_match() is None:
            pytest.skip("oracle/_ref/libmatch_shim.so not built (needs the product library)")
    return SideBySide(oracle)


def ran(side, fn, *args):
    before = side.calls
    fn(side, *args)
    assert side.calls > before          # the scene did go through both libraries


@pytest.mark.parametrize("seed,nnratio,ori", [(5, 0.9, True), (6, 0.7, False)])
def test_search_for_initialization(side, seed, nnratio, ori):
    ran(side, T.test_search_for_initialization, seed, nnratio, ori)


@pytest.mark.parametrize("seed,th,nnratio", [(7, 1.0, 0.8), (8, 3.0, 0.8), (9, 5.0, 0.6)])
def test_search_by_projection_track(side, seed, th, nnratio):
    ran(side, T.test_search_by_projection_track, seed, th, nnratio)


@pytest.mark.parametrize("seed,th", [(20, 3.0), (21, 5.0)])
def test_fuse(side, seed, th):
    ran(side, T.test_fuse, seed, th)


@pytest.mark.parametrize("seed,th,scale", [(22, 4.0, 2.0), (23, 3.0, 0.5)])
def test_fuse_sim3(side, seed, th, scale):
    ran(side, T.test_fuse_sim3, seed, th, scale)


@pytest.mark.parametrize("seed,scale", [(24, 2.0), (25, 1.0)])
def test_search_by_projection_sim3(side, seed, scale):
    ran(side, T.test_search_by_projection_sim3, seed, scale)


def test_search_by_sim3(side):
    ran(side, T.test_search_by_sim3)


@pytest.mark.parametrize("seed,th,ori", [(30, 7.0, True), (31, 15.0, False)])
def test_search_by_projection_last_frame(side, seed, th, ori):
    ran(side, T.test_search_by_projection_last_frame, seed, th, ori)


@pytest.mark.parametrize("seed,th,orb_dist,ori", [(32, 10.0, 100, True), (33, 3.0, 64, False)])
def test_search_by_projection_relocalisation(side, seed, th, orb_dist, ori):
    ran(side, T.test_search_by_projection_relocalisation, seed, th, orb_dist, ori)


@pytest.mark.parametrize("seed,nnratio,ori", [(0, 0.7, True), (1, 0.9, False), (2, 0.6, True)])
def test_search_by_bow(side, seed, nnratio, ori):
    ran(side, T.test_search_by_bow, seed, nnratio, ori)


@pytest.mark.parametrize("seed,ori", [(3, False), (4, True)])
def test_search_for_triangulation(side, seed, ori):
    ran(side, T.test_search_for_triangulation, seed, ori)


def test_descriptor_distance(side):
    rng = np.random.default_rng(0)
    with side._o.matcher_side("shim"):
        L = side._o.ref_match()
        for _ in range(50):
            a, b = rng.integers(0, 256, 32, dtype=np.uint8), rng.integers(0, 256, 32, dtype=np.uint8)
            assert L.ref_descriptor_distance(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)) == int(np.unpackbits(a ^ b).sum())


def test_tie_storm_scenes(side):
    """the same through both implementations with ties everywhere (six descriptor patterns): the visiting order alone decides"""
    T.TIES = True
    try:
        ran(side, T.test_fuse, 40, 4.0)
        ran(side, T.test_fuse_sim3, 41, 4.0, 2.0)
        ran(side, T.test_search_by_projection_sim3, 42, 2.0)
        ran(side, T.test_search_by_projection_last_frame, 43, 7.0, True)
        ran(side, T.test_search_by_projection_relocalisation, 44, 10.0, 100, False)
    finally:
        T.TIES = False
    ran(side, T.test_ties_track_and_initialization)


def test_shims_type_check_against_the_reference_headers(oracle):
    """shim/ORBextractor_shim.cpp and the two matcher shims define members of the reference's classes: the compiler checks every
    signature against cslam/ORBextractor.h / cslam/ORBmatcher.h as they are in the reference tree (OpenCV and the Frame / KeyFrame /
    MapPoint classes are the stand-ins of oracle/ref_stub)."""
    import os
    import subprocess
    if not os.path.isdir("/root/reference/cslam/include"):
        pytest.skip("reference tree absent")
    subprocess.check_call(["make", "-C", os.path.dirname(oracle.__file__), "-s", "shim-check"])
