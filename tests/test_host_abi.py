"""CPU-side checks of the product: the C-ABI library loads, exports every symbol include/ccm_b200.h declares, refuses to
compute without a device (no CPU fallback), and the host-only helpers agree with the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

from ccm_slam_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "ccm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ccm_[a-z0-9_]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported():
    lib = api.lib()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ccm_b200.h but not exported by libccm_b200.so"


def test_no_cpu_fallback_without_a_device():
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.CCMError) as e:
        api.init(0)
    assert e.value.code == -2
    with pytest.raises(api.CCMError) as e:
        api.ba_solve(synth.make_config("tiny"), iterations=1)
    assert e.value.code == -2
    with pytest.raises(api.CCMError):
        api.hamming_matrix(np.zeros((2, 32), np.uint8), np.zeros((2, 32), np.uint8))


def test_pose_conversion_helpers_match_oracle(oracle):
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(0)
    T = np.tile(np.eye(4, dtype=np.float32), (64, 1, 1))
    rv = rng.normal(size=(64, 3)); rv *= (np.pi * rng.uniform(0, 1, size=(64, 1))) / np.linalg.norm(rv, axis=1, keepdims=True)
    T[:, :3, :3] = Rotation.from_rotvec(rv).as_matrix(); T[:, :3, 3] = rng.normal(size=(64, 3))
    qt = api.poses_from_Tcw_f32(T)
    for i in range(64):
        assert np.array_equal(qt[i], oracle.pose_from_Tcw_f32(T[i]))
        assert np.array_equal(api.poses_to_Tcw_f32(qt[i])[0], oracle.pose_to_Tcw_f32(qt[i]))


def test_product_does_not_reference_the_oracle():
    """The product path (package + csrc) must never import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "ccm_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "oracle/" not in txt.replace("the oracle", ""), f
    out = os.popen(f"ldd {api.LIB_PATH}").read()
    assert "liboracle" not in out
