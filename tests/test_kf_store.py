"""the wire format of ccmslam_msgs::KF -> device-resident descriptors (ccm_kfstore_*).

CPU half: the wire decode (Converter::fromCvKeyPointMsg over the packed 15-byte ROS records) against a restatement with struct.unpack.
GPU half: a keyframe put once is read back bit for bit; Hamming matrices, SearchByBoW(kf, kf) and the BoW transform over resident
operands equal the host-operand entry points (which are index-exact against the oracle, tests/test_gpu_frontend.py); descriptors cross
the bus once per keyframe, whatever the number of matcher calls."""
import struct

import numpy as np
import pytest

from ccm_slam_b200 import api
from ccm_slam_b200 import frontend as fe


def _kps(rng, n):
    k = np.zeros(n, fe.KP_DTYPE)
    k["x"] = rng.uniform(0, 752, n).astype(np.float32); k["y"] = rng.uniform(0, 480, n).astype(np.float32)
    k["size"] = (31 * 1.2 ** rng.integers(0, 8, n)).astype(np.float32); k["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    k["response"] = rng.integers(7, 200, n).astype(np.float32); k["octave"] = rng.integers(0, 8, n)
    return k


def test_wire_keypoints_roundtrip_like_the_reference():
    """toCvKeyPointMsg truncates size / response to uint8 and octave to int8 (S/Converter.cc:166-178); fromCvKeyPointMsg widens them
    back (:180-192).  The decode runs on the host and needs no device."""
    rng = np.random.default_rng(0)
    k = _kps(rng, 257)
    k["size"][:3] = [31.0, 37.2, 255.9]; k["octave"][:2] = [0, 7]
    w = fe.wire_keypoints(k)
    raw = w.tobytes()
    assert len(raw) == 15 * len(k)
    got = fe.wire_keypoints_decode(w)
    for i in (0, 1, 2, 100, 256):
        x, y, size, angle, resp, octv = struct.unpack_from("<ffBfBb", raw, 15 * i)
        assert (got["x"][i], got["y"][i], got["angle"][i]) == (np.float32(x), np.float32(y), np.float32(angle))
        assert got["size"][i] == float(size) and got["response"][i] == float(resp) and got["octave"][i] == octv
    assert np.array_equal(got["x"], k["x"]) and np.array_equal(got["angle"], k["angle"]) and np.array_equal(got["octave"], k["octave"])
    assert np.array_equal(got["size"], np.floor(k["size"])) and np.array_equal(got["response"], np.floor(k["response"]))
    assert len(fe.wire_keypoints_decode(w[:0])) == 0


def test_store_needs_a_device():
    if api.device_count() > 0:
        pytest.skip("a device is present")
    with pytest.raises(api.CCMError):
        fe.KeyFrameStore()                                   # no CPU fallback: CCM_ERR_NO_DEVICE


@pytest.mark.gpu
def test_resident_keyframes_match_the_host_operand_paths(oracle):
    from ccm_slam_b200 import synth_match as sm
    api.init(0)
    rng = np.random.default_rng(3)
    st = fe.KeyFrameStore()
    n1, n2 = 987, 1013
    voc = sm.make_vocabulary(k=6, L=3, seed=31)
    d1 = sm.make_voc_features(voc, n=n1, seed=32); d2 = d1[rng.permutation(n1)][: n1 - 100].copy()
    d2 = np.vstack([d2, sm.make_voc_features(voc, n=n2 - len(d2), seed=33)])
    flip = rng.integers(0, 256, size=(len(d2), 2)); d2[np.arange(len(d2)), flip[:, 0] % 32] ^= (1 << (flip[:, 1] % 8)).astype(np.uint8)
    k1, k2 = _kps(rng, n1), _kps(rng, n2)
    out1 = st.put_wire(11, fe.wire_keypoints(k1), d1)
    st.put_wire(12, fe.wire_keypoints(k2), d2)
    assert np.array_equal(out1["angle"], k1["angle"]) and st.features(11) == n1 and st.features(12) == n2 and st.keyframes() == 2
    gk, gd = st.get(11)
    assert np.array_equal(gd, d1) and np.array_equal(gk["octave"], k1["octave"])
    bytes_after_ingest = st.h2d_bytes()
    assert bytes_after_ingest == 32 * (n1 + n2)
    # distances: both operands resident / host queries against a resident keyframe
    assert np.array_equal(st.hamming(11, 12), api.hamming_matrix(d1, d2))
    Q = d2[5:300]
    assert np.array_equal(st.hamming_query(Q, 11), api.hamming_matrix(Q, d1))
    # BoW transform over the resident copy == over host descriptors == the oracle
    V = fe.ORBVocabulary(voc); R = oracle.Vocabulary(voc)
    t1, t2 = st.transform(11, V, 1), st.transform(12, V, 1)
    h1 = V.transform(d1, 1); r1 = R.transform(d1, 1)
    for key in ("word", "node", "bow_id", "bow_val", "fv_node_id", "fv_node_ptr", "fv_feat"):
        assert np.array_equal(t1[key], h1[key]), key
    assert np.array_equal(t1["word"], r1["word"]) and np.array_equal(t1["bow_val"], r1["bow_val"]) and np.array_equal(t1["fv_feat"], r1["fv_feat"])
    # SearchByBoW(kf, kf) on resident operands == host operands == oracle
    fv1, fv2 = fe.FeatureVector(t1["node"]), fe.FeatureVector(t2["node"])
    ofv1, ofv2 = oracle.FeatureVector(t1["node"]), oracle.FeatureVector(t2["node"])
    has1 = (rng.random(n1) < 0.7).astype(np.uint8); has2 = (rng.random(n2) < 0.7).astype(np.uint8)
    a1, a2 = fe.wire_keypoints_decode(fe.wire_keypoints(k1))["angle"], fe.wire_keypoints_decode(fe.wire_keypoints(k2))["angle"]
    for nnratio, ori in ((0.8, True), (0.9, False)):
        got, n = st.SearchByBoW_KF_KF(11, has1, fv1, 12, has2, fv2, nnratio, ori)
        host, hn = fe.ORBmatcher(nnratio, ori).SearchByBoW_KF_KF(d1, has1, a1, fv1, d2, has2, a2, fv2)
        ref, rn = oracle.match_bow_kf_kf(d1, has1, a1, ofv1, d2, has2, a2, ofv2, nnratio, ori)
        assert n == hn == rn and np.array_equal(got, host) and np.array_equal(got, ref)
    assert n > 20
    assert st.h2d_bytes() == bytes_after_ingest              # no descriptor crossed the bus again
    # erase / reuse / update in place / unknown ids
    st.erase(11); st.erase(11)
    assert st.features(11) == -1 and st.keyframes() == 1
    with pytest.raises(api.CCMError):
        st.hamming(11, 12)
    d3 = rng.integers(0, 256, size=(n1, 32), dtype=np.uint8)
    st.put(13, k1, d3)                                        # same size: takes the freed range
    assert np.array_equal(st.get(13)[1], d3) and np.array_equal(st.get(12)[1], d2)
    st.put(12, k2[:500], d2[:500])                            # an update message with another N
    assert st.features(12) == 500 and np.array_equal(st.hamming(13, 12), api.hamming_matrix(d3, d2[:500]))
    st.put(14, k1[:0], d1[:0])
    assert st.features(14) == 0 and st.hamming(14, 12).shape == (0, 500)
    V.close(); R.close(); st.close()
