"""CPU suite: shim/ORBVocabulary_shim.cpp as a drop-in for the two ComputeBoW bodies.

The shim (Frame::ComputeBoW, KeyFrame::ComputeBoW, ccm_b200_load_vocabulary's reader of the vocabulary text file) is compiled against the
reference's own cslam/ORBVocabulary.h and DBoW2 headers and run next to the reference's own
`mpORBvocabulary->transform(toDescriptorVector(mDescriptors), mBowVec, mFeatVec, 4)` on the same vocabulary file and descriptors
(oracle/ref_voc_shim_wrap.cpp -> oracle/_ref/libvoc_shim.so).  The device descent is doubled by the CPU oracle, the containers come from
the library's own host half ccm_bow_assemble (oracle/ccm_voc_double.cpp).  BowVector and FeatureVector must be identical — ids, f64
values bit for bit, feature lists.  Skipped where the reference tree or the product library is absent."""
import ctypes as C
import os

import numpy as np
import pytest

from ccm_slam_b200 import synth_match as sm

SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref", "libvoc_shim.so")


@pytest.fixture(scope="module")
def vlib(oracle):
    if oracle.build_ref() is None or not os.path.exists(SO):
        pytest.skip("oracle/_ref/libvoc_shim.so not available")
    L = C.CDLL(SO)
    L.vshim_load.restype = C.c_void_p
    return L


def compute(L, h, side, desc):
    n = len(desc)
    d = np.ascontiguousarray(desc, np.uint8)
    bid = np.zeros(max(n, 1), np.uint32); bval = np.zeros(max(n, 1)); fid = np.zeros(max(n, 1), np.uint32); fptr = np.zeros(n + 1, np.int32)
    ff = np.zeros(max(n, 1), np.uint32); bn = C.c_int32(0); fn = C.c_int32(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.vshim_compute_bow(C.c_void_p(h), side, p(d), n, p(bid), p(bval), C.byref(bn), p(fid), p(fptr), p(ff), C.byref(fn)) == 0
    return bid[:bn.value].copy(), bval[:bn.value].copy(), fid[:fn.value].copy(), fptr[:fn.value + 1].copy(), ff[:fptr[fn.value]].copy()


@pytest.mark.parametrize("k,L_,scoring,weighting", [(10, 3, 0, 0), (6, 4, 0, 0), (8, 3, 1, 1), (5, 3, 5, 2), (9, 2, 2, 3)])
def test_compute_bow_matches_reference_transform(vlib, oracle, tmp_path, k, L_, scoring, weighting):
    voc = sm.make_vocabulary(k=k, L=L_, seed=k + L_, scoring=scoring, weighting=weighting)
    path = str(tmp_path / "voc.txt")
    oracle.write_vocabulary_text(voc, path)
    h = vlib.vshim_load(path.encode())
    assert h
    for n, seed in ((1200, 1), (7, 2), (0, 3)):
        feat = sm.make_voc_features(voc, n=n, seed=seed) if n else np.zeros((0, 32), np.uint8)
        want = compute(vlib, h, 2, feat)
        for side in (0, 1):
            got = compute(vlib, h, side, feat)
            for a, b in zip(got, want):
                assert np.array_equal(a, b)
        if n > 100:
            assert len(want[0]) > 20 and len(want[2]) >= 1
    vlib.vshim_free(C.c_void_p(h))
