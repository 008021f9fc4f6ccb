"""CPU suite: shim/MapUpdate_shim.cpp (cslam::UpdateMapAfterGBA) at the class boundary, on stand-in Map / KeyFrame / MapPoint objects,
next to a restatement of the loop it replaces (Map::RunGBA, S/Map.cpp:1441-1570 = MapMerger::RunGBA, S/MapMerger.cpp:637-753) —
oracle/ref_map_update_wrap.cpp; ccm_gba_map_update is doubled by the oracle in that library.

 * which keyframes and points are touched, how often their setters are called, which flags change: exact;
 * values: the restated loop uses the stand-in cv::Mat, whose small products round as cv::gemm's do (f32, left to right — pinned against
   cv2 4.13 in tests/test_map_update.py), the shim the oracle: bit for bit;
 * the shim on objects against the oracle on the flat arrays of the same scene: bit for bit (flattening and write-back lose nothing)."""
import ctypes as C
import os

import numpy as np
import pytest

from ccm_slam_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "..", "oracle", "_ref", "libmap_update_shim.so")


@pytest.fixture(scope="module")
def mapw(oracle):
    if not os.path.exists(SO):
        if not os.path.isdir("/root/reference"):
            pytest.skip("oracle/_ref/libmap_update_shim.so not built (needs the reference's headers)")
        oracle.build_ref()
    return C.CDLL(SO)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def run(mapw, sc, mode):
    K = len(sc["kf_parent"]); P = len(sc["mp_state"])
    a = [np.ascontiguousarray(sc["kf_parent"], np.int32), np.ascontiguousarray(sc["kf_optimized"], np.uint8), np.ascontiguousarray(sc["kf_Tcw"], np.float32),
         np.ascontiguousarray(sc["kf_TcwGBA"], np.float32), np.ascontiguousarray(sc["mp_state"], np.uint8), np.ascontiguousarray(sc["mp_ref"], np.int32),
         np.ascontiguousarray(sc["mp_pos"], np.float32), np.ascontiguousarray(sc["mp_pos_gba"], np.float32)]
    o = dict(pose=np.zeros((max(K, 1), 4, 4), np.float32), bef=np.zeros((max(K, 1), 4, 4), np.float32), gba=np.zeros((max(K, 1), 4, 4), np.float32),
             kf_info=np.zeros((max(K, 1), 3), np.int32), pos=np.zeros((max(P, 1), 3), np.float32), mp_info=np.zeros((max(P, 1), 2), np.int32))
    rc = mapw.mapw_update(mode, K, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), P, _p(a[4]), _p(a[5]), _p(a[6]), _p(a[7]),
                          _p(o["pose"]), _p(o["bef"]), _p(o["gba"]), _p(o["kf_info"]), _p(o["pos"]), _p(o["mp_info"]))
    assert rc == 0
    return {k: (v[:K] if k in ("pose", "bef", "gba", "kf_info") else v[:P]) for k, v in o.items()}


CASES = [dict(K=200, P=5000, seed=0), dict(K=1, P=50, seed=1, n_origins=1), dict(K=800, P=6000, seed=2, chain=1.0, n_origins=1, new_kf_frac=0.3),
         dict(K=300, P=0, seed=3), dict(K=64, P=3000, seed=4, chain=0.0, n_origins=4, outside_frac=0.2), dict(K=500, P=4000, seed=5, new_kf_frac=0.0)]


@pytest.mark.parametrize("kw", CASES, ids=lambda kw: "K%d-P%d-s%d" % (kw["K"], kw["P"], kw["seed"]))
def test_shim_next_to_the_restated_loop(mapw, kw):
    sc = synth.make_map_update(**kw)
    ref = run(mapw, sc, 0); shim = run(mapw, sc, 1)
    assert np.array_equal(ref["kf_info"], shim["kf_info"]) and np.array_equal(ref["mp_info"], shim["mp_info"])      # flags and setter call counts
    for key in ("pose", "bef", "gba", "pos"):                                                                      # every value: bit for bit
        assert np.array_equal(ref[key], shim[key], equal_nan=True), key
    touched = ref["kf_info"][:, 0] == 1
    assert touched.sum() == (sc["kf_parent"] != -2).sum()
    assert (ref["kf_info"][touched, 1] == 1).all() and (ref["kf_info"][~touched, 1] == 0).all()                    # SetPose exactly once per visited keyframe
    assert np.array_equal(ref["bef"][touched], sc["kf_Tcw"][touched])                                              # mTcwBefGBA = the pose before
    assert np.array_equal(shim["pose"][touched], shim["gba"][touched])                                             # SetPose(mTcwGBA)


@pytest.mark.parametrize("kw", CASES, ids=lambda kw: "K%d-P%d-s%d" % (kw["K"], kw["P"], kw["seed"]))
def test_shim_loses_nothing_around_the_flat_call(oracle, mapw, kw):
    sc = synth.make_map_update(**kw)
    shim = run(mapw, sc, 1); flat = oracle.gba_map_update(sc)
    vis = flat["kf_visited"].astype(bool)
    assert np.array_equal(shim["kf_info"][:, 0].astype(bool), vis)
    assert np.array_equal(shim["pose"][vis], flat["kf_TcwGBA"][vis]) and np.array_equal(shim["pose"][~vis], sc["kf_Tcw"][~vis])
    corr = flat["mp_corrected"].astype(bool)
    assert np.array_equal(shim["mp_info"][:, 0].astype(bool), corr) and np.array_equal(shim["mp_info"][:, 1], corr.astype(np.int32))
    assert np.array_equal(shim["pos"], flat["mp_pos"])


def test_origin_without_ba_result_is_an_error(mapw):
    sc = synth.make_map_update(K=20, P=10, seed=3)
    sc["kf_optimized"][0] = 0
    K, P = 20, 10
    o = [np.zeros((K, 4, 4), np.float32) for _ in range(3)] + [np.zeros((K, 3), np.int32), np.zeros((P, 3), np.float32), np.zeros((P, 2), np.int32)]
    a = [np.ascontiguousarray(sc[k], t) for k, t in (("kf_parent", np.int32), ("kf_optimized", np.uint8), ("kf_Tcw", np.float32), ("kf_TcwGBA", np.float32),
                                                     ("mp_state", np.uint8), ("mp_ref", np.int32), ("mp_pos", np.float32), ("mp_pos_gba", np.float32))]
    rc = mapw.mapw_update(1, K, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), P, _p(a[4]), _p(a[5]), _p(a[6]), _p(a[7]), *[_p(x) for x in o])
    assert rc == 1          # the shim raises (the reference would multiply by an empty Mat)
