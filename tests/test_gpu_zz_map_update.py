"""GPU suite (first device run in round 2: profiles/r2/map_update_gpu.log): the map update
after a global BA through the C ABI (ccm_gba_map_update: host keyframe pass + k_map_update_points) against the CPU oracle — the same
f32 operations in the same order on both sides, so bit for bit.  The arithmetic has run on the host (tests/test_map_update.py)."""
import os

import numpy as np
import pytest

from ccm_slam_b200 import api, synth

pytestmark = pytest.mark.gpu   # first device run: round 2, profiles/r2/map_update_gpu.log


@pytest.mark.parametrize("kw", [dict(K=200, P=5000, seed=0), dict(K=1, P=50, seed=1, n_origins=1), dict(K=300, P=0, seed=3),
                                dict(K=2000, P=20000, seed=2, chain=1.0, n_origins=1, new_kf_frac=0.3), dict(K=10000, P=1000000, seed=7)])
def test_map_update_matches_the_oracle(oracle, kw):
    assert api.device_count() > 0
    api.init(0)
    sc = synth.make_map_update(**kw)
    l0 = api.kernel_launches()
    got = api.gba_map_update(sc); ref = oracle.gba_map_update(sc)
    assert api.kernel_launches() == l0 + (1 if kw["P"] else 0)
    vis = ref["kf_visited"].astype(bool)
    assert np.array_equal(got["kf_visited"], ref["kf_visited"]) and np.array_equal(got["mp_corrected"], ref["mp_corrected"])
    assert np.array_equal(got["kf_TcwGBA"][vis], ref["kf_TcwGBA"][vis])
    assert np.array_equal(got["mp_pos"], ref["mp_pos"], equal_nan=True)
