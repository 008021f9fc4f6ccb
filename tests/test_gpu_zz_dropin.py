"""GPU suite (runs last): the reference-side shims as a drop-in for cslam::ORBmatcher, on the device.

Same construction as tests/test_shim_dropin.py — shim/ORBmatcher_shim.cpp + shim/ORBmatcher_proj_shim.cpp behind the C wrappers of
oracle/ref_match_wrap.cpp, next to the reference's own ORBmatcher.cpp — but linked against the product library itself
(oracle/_ref/libmatch_shim_gpu.so): the Hamming matrices come from k_hamming on the GPU.  Every scene goes through both
implementations of the class and must give identical results.  All eleven methods have been run this way on the CPU with the device half
doubled (tests/test_shim_dropin.py), and the optimiser shim against the reference's own Optimizer.cpp (tests/test_shim_optimizer.py)."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import test_oracle_vs_reference_matchers as T
from tests.test_shim_dropin import same

# Loads a second native library (the shims linked against the product) into the test process.
pytestmark = pytest.mark.gpu   # first device run: round 2, profiles/r2/dropin_gpu.log


class SideBySideGPU:
    def __init__(self, oracle, skip=()):
        self._o, self.calls, self.skip = oracle, 0, set(skip)

    def __getattr__(self, name):
        f = getattr(self._o, name)
        if not name.startswith("ref_") or name in self.skip or name == "ref_match" or not callable(f):
            return f

        def both(*a, **k):
            r = f(*a, **k)
            with self._o.matcher_side("shim_gpu"):
                s = f(*a, **k)
            same(r, s)
            self.calls += 1
            return r
        return both


@pytest.fixture(scope="module")
def side(oracle):
    from ccm_slam_b200 import api
    if api.device_count() == 0:
        pytest.skip("no CUDA device")
    api.init(0)
    if oracle.ref_match() is None:
        pytest.skip("no oracle/_ref/libmatch_ref.so")
    with oracle.matcher_side("shim_gpu"):
        if oracle.ref_match() is None:
            pytest.skip("no oracle/_ref/libmatch_shim_gpu.so")
    return SideBySideGPU(oracle)


def ran(side, fn, *args):
    before = side.calls
    fn(side, *args)
    assert side.calls > before


def test_projection_guided_methods(side):
    ran(side, T.test_search_for_initialization, 5, 0.9, True)
    ran(side, T.test_search_by_projection_track, 8, 3.0, 0.8)
    ran(side, T.test_fuse, 20, 3.0)
    ran(side, T.test_fuse_sim3, 22, 4.0, 2.0)
    ran(side, T.test_search_by_projection_sim3, 24, 2.0)
    ran(side, T.test_search_by_sim3)
    ran(side, T.test_search_by_projection_last_frame, 30, 7.0, True)
    ran(side, T.test_search_by_projection_relocalisation, 33, 3.0, 64, False)
    T.TIES = True                                   # and with ties everywhere: the visiting order alone decides
    try:
        ran(side, T.test_fuse, 40, 4.0)
        ran(side, T.test_search_by_projection_sim3, 42, 2.0)
        ran(side, T.test_search_by_projection_last_frame, 43, 7.0, True)
    finally:
        T.TIES = False
    ran(side, T.test_ties_track_and_initialization)


def test_bow_and_triangulation_methods(side):
    ran(side, T.test_search_by_bow, 0, 0.7, True)
    ran(side, T.test_search_by_bow, 1, 0.9, False)
    ran(side, T.test_search_for_triangulation, 3, False)
    ran(side, T.test_search_for_triangulation, 4, True)
    rng = np.random.default_rng(0)
    with side._o.matcher_side("shim_gpu"):
        for _ in range(20):
            a, b = rng.integers(0, 256, 32, dtype=np.uint8), rng.integers(0, 256, 32, dtype=np.uint8)
            assert side._o.ref_match().ref_descriptor_distance(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)) == int(np.unpackbits(a ^ b).sum())


def test_optimizer_shim_on_the_device(oracle):
    """shim/Optimizer_shim.cpp through cslam::Optimizer's interface with the real device entry points underneath: MapFusionGBA on a
    stand-in map must land where the reference's own Optimizer.cpp lands, within the device path's tolerance (f32 poses / points)."""
    from ccm_slam_b200 import api, synth
    from tests import shim_optimizer_harness as H
    if api.device_count() == 0:
        pytest.skip("no CUDA device")
    api.init(0)
    p = synth.make_config("small")
    sc = H.scene_from_problem(p, oracle, seed=3, map_id=0, bad_kf=0.1, bad_mp=0.1)
    sc["kf_bad"][0] = 0
    H.use_device(False); H.use_reference(True)                     # the CPU side: the reference's own Optimizer.cpp (oracle/_ref/liboptimizer_ref.so)
    try:
        if H.lib() is None:
            pytest.skip("no oracle/_ref/liboptimizer_ref.so")
        cpu = H.run_gba(sc, 0, 8, True, (0, 0))
    finally:
        H.use_reference(False)
    H.use_device(True)
    try:
        if H.lib() is None:
            pytest.skip("no oracle/_ref/liboptimizer_shim_gpu.so")
        l0 = api.kernel_launches()
        gpu = H.run_gba(sc, 0, 8, True, (0, 0))
        assert api.kernel_launches() > l0
    finally:
        H.use_device(False)
    assert np.array_equal(gpu["kf_set_pose"], cpu["kf_set_pose"]) and np.array_equal(gpu["mp_update_normal"], cpu["mp_update_normal"])
    assert np.abs(gpu["kf_Tcw"] - cpu["kf_Tcw"]).max() <= 1e-4 * max(1.0, np.abs(cpu["kf_Tcw"]).max())
    assert np.abs(gpu["mp_pos"] - cpu["mp_pos"]).max() <= 1e-4 * np.abs(cpu["mp_pos"]).max()
