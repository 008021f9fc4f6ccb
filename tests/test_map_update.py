"""CPU suite, SURVEY.md §8(f) rank 1: the map update that follows a global BA (Map::RunGBA, S/Map.cpp:1441-1570 =
MapMerger::RunGBA, S/MapMerger.cpp:637-753) — spanning-tree propagation of mTcwGBA, correction of every map point.

 * the oracle (oracle/map_update_oracle.cpp) against OpenCV's own arithmetic — the loop evaluated with cv2.gemm 4.13, as a committed
   fixture (tests/golden/map_update_cv2.npz + its generator) and live where cv2 imports: bit for bit;
 * the oracle against an f64 numpy evaluation of the same rules: a few f32 ulps;
 * the product's arithmetic (csrc/map_update_math.cuh: the host keyframe pass as shipped, the kernel body as a plain loop;
   tests/host/map_update_host.cpp, g++) against the oracle: bit for bit;
 * which keyframes / points are touched at all: exact.
The kernel launch itself is tests/test_gpu_zz_dropin.py (opt-in until it has run on a device)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from ccm_slam_b200 import api, synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("mu") / "libmap_update_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-fno-fast-math", "-ffp-contract=off", "-o", so,
                           os.path.join(HERE, "host", "map_update_host.cpp")])
    return C.CDLL(so).mu_host_update


def f64_update(sc):
    """the rules in f64, parents before children (the generator numbers keyframes that way)"""
    K = len(sc["kf_parent"]); T = sc["kf_Tcw"].astype(np.float64); G = sc["kf_TcwGBA"].astype(np.float64).copy()
    vis = np.zeros(K, bool)
    for k in range(K):
        p = sc["kf_parent"][k]
        if p == -1:
            vis[k] = True
        elif p >= 0 and vis[p]:
            vis[k] = True
            if not sc["kf_optimized"][k]:
                G[k] = (T[k] @ np.linalg.inv(T[p])) @ G[p]
    out = sc["mp_pos"].astype(np.float64).copy(); corr = np.zeros(len(out), bool)
    for i in range(len(out)):
        st = sc["mp_state"][i]
        if st == 1:
            out[i] = sc["mp_pos_gba"][i]; corr[i] = True
        elif st == 2 and sc["mp_ref"][i] >= 0 and vis[sc["mp_ref"][i]]:
            r = sc["mp_ref"][i]
            xc = T[r, :3, :3] @ out[i] + T[r, :3, 3]
            Twc = np.linalg.inv(G[r])
            out[i] = Twc[:3, :3] @ xc + Twc[:3, 3]; corr[i] = True
    return G, vis, out, corr


CASES = [dict(K=200, P=5000, seed=0), dict(K=1, P=50, seed=1, n_origins=1), dict(K=2000, P=20000, seed=2, chain=1.0, n_origins=1, new_kf_frac=0.3),
         dict(K=300, P=0, seed=3), dict(K=64, P=3000, seed=4, chain=0.0, n_origins=4, outside_frac=0.2), dict(K=500, P=4000, seed=5, new_kf_frac=0.0),
         dict(K=40, P=1000, seed=6, new_kf_frac=0.9)]


@pytest.mark.parametrize("kw", CASES, ids=lambda kw: "K%d-P%d-s%d" % (kw["K"], kw["P"], kw["seed"]))
def test_oracle_against_f64(oracle, kw):
    sc = synth.make_map_update(**kw)
    r = oracle.gba_map_update(sc)
    G, vis, out, corr = f64_update(sc)
    assert np.array_equal(r["kf_visited"].astype(bool), vis) and np.array_equal(r["mp_corrected"].astype(bool), corr)
    got = r["kf_TcwGBA"][vis].astype(np.float64); want = G[vis]
    assert not np.isnan(got).any()
    depth = 1 if kw.get("chain", 0.7) < 1.0 else int((~sc["kf_optimized"].astype(bool)).sum())      # roundings pile up along a run of new keyframes
    scale = np.abs(want).max(axis=(1, 2), keepdims=True)
    assert (np.abs(got - want) <= 64 * np.finfo(np.float32).eps * scale * max(1, min(depth, 40))).all()
    o = r["mp_pos"][corr].astype(np.float64)
    assert not np.isnan(o).any()
    # a corrected point goes through two poses with translations of the order of 10: absolute error of a few f32 ulps of that scale
    assert np.abs(o - out[corr]).max(initial=0.0) <= 2e-3 * max(1, min(depth, 40))
    assert np.array_equal(r["mp_pos"][~corr], sc["mp_pos"][~corr])                                    # untouched points come back as they went in
    assert np.array_equal(r["mp_pos"][sc["mp_state"] == 1], sc["mp_pos_gba"][sc["mp_state"] == 1])


@pytest.mark.parametrize("kw", CASES, ids=lambda kw: "K%d-P%d-s%d" % (kw["K"], kw["P"], kw["seed"]))
def test_product_arithmetic_on_host_is_the_oracles(oracle, host, kw):
    sc = synth.make_map_update(**kw)
    a = oracle.gba_map_update(sc); b = oracle.gba_map_update(sc, fn=host)
    vis = a["kf_visited"].astype(bool)
    assert np.array_equal(a["kf_visited"], b["kf_visited"]) and np.array_equal(a["mp_corrected"], b["mp_corrected"])
    assert np.array_equal(a["kf_TcwGBA"][vis], b["kf_TcwGBA"][vis])
    assert np.array_equal(a["mp_pos"], b["mp_pos"], equal_nan=True)
    assert vis.sum() > 0


def test_new_keyframes_follow_their_parent(oracle):
    """a keyframe the BA did not hold keeps its pose RELATIVE to its parent: Tcw_child * Twc_parent is the same before and after"""
    sc = synth.make_map_update(K=400, P=0, seed=11, new_kf_frac=0.4)
    r = oracle.gba_map_update(sc)
    T = sc["kf_Tcw"].astype(np.float64); G = r["kf_TcwGBA"].astype(np.float64)
    n = 0
    for k in np.flatnonzero((sc["kf_optimized"] == 0) & (r["kf_visited"] == 1)):
        p = sc["kf_parent"][k]
        assert np.abs(T[k] @ np.linalg.inv(T[p]) - G[k] @ np.linalg.inv(G[p])).max() < 2e-4
        n += 1
    assert n > 50


def test_points_keep_their_place_in_the_reference_camera(oracle):
    sc = synth.make_map_update(K=100, P=4000, seed=12)
    r = oracle.gba_map_update(sc)
    sel = np.flatnonzero((sc["mp_state"] == 2) & (r["mp_corrected"] == 1))
    assert len(sel) > 200
    k = sc["mp_ref"][sel]
    T = sc["kf_Tcw"].astype(np.float64)[k]; G = r["kf_TcwGBA"].astype(np.float64)[k]
    before = np.einsum("nij,nj->ni", T[:, :3, :3], sc["mp_pos"][sel].astype(np.float64)) + T[:, :3, 3]
    after = np.einsum("nij,nj->ni", G[:, :3, :3], r["mp_pos"][sel].astype(np.float64)) + G[:, :3, 3]
    assert np.abs(before - after).max() < 5e-4


def test_conventions(oracle, host):
    sc = synth.make_map_update(K=30, P=100, seed=13)
    bad = dict(sc); bad["kf_optimized"] = sc["kf_optimized"].copy(); bad["kf_optimized"][0] = 0      # an origin without a BA result
    with pytest.raises(ValueError):
        oracle.gba_map_update(bad)
    with pytest.raises(ValueError):
        oracle.gba_map_update(bad, fn=host)
    # a point whose reference keyframe holds a BA result but hangs outside the tree stays where it is (mTcwBefGBA was never set there)
    sc2 = synth.make_map_update(K=60, P=2000, seed=14, outside_frac=0.3)
    r = oracle.gba_map_update(sc2)
    out_kf = (sc2["kf_parent"] == -2)
    assert out_kf.sum() > 5 and (r["kf_visited"][out_kf] == 0).all()
    via_out = (sc2["mp_state"] == 2) & (sc2["mp_ref"] >= 0) & out_kf[np.maximum(sc2["mp_ref"], 0)]
    assert via_out.sum() > 20 and (r["mp_corrected"][via_out] == 0).all()
    # empty map
    e = oracle.gba_map_update(synth.make_map_update(K=0, P=0, seed=1, n_origins=0))
    assert len(e["kf_visited"]) == 0 and len(e["mp_corrected"]) == 0


def test_library_entry_point_needs_a_device():
    """the C ABI entry point exists and refuses to run without a GPU (no CPU fallback behind it)"""
    assert hasattr(api.lib(), "ccm_gba_map_update")
    if api.device_count() == 0:
        with pytest.raises(api.CCMError):
            api.gba_map_update(synth.make_map_update(K=10, P=10, seed=1))


# ---- the pin: OpenCV's own arithmetic -------------------------------------------------------------------------------------------
def _golden_cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_map_update_golden", os.path.join(HERE, "golden", "make_map_update_golden.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    return mod


def _same_as(r, w):
    vis = w["kf_visited"].astype(bool)
    assert np.array_equal(r["kf_visited"], w["kf_visited"]) and np.array_equal(r["mp_corrected"], w["mp_corrected"])
    assert np.array_equal(r["kf_TcwGBA"][vis], w["kf_TcwGBA"][vis])
    assert np.array_equal(r["mp_pos"], w["mp_pos"], equal_nan=True)


def test_oracle_reproduces_the_opencv_fixture(oracle, host):
    """tests/golden/map_update_cv2.npz: the loop evaluated with cv2.gemm (4.13) for every cv::Mat product — bit for bit, for the oracle
    and for the product's arithmetic run on the host"""
    mod = _golden_cases()
    z = np.load(os.path.join(HERE, "golden", "map_update_cv2.npz"))
    for n, kw in enumerate(mod.CASES):
        sc = synth.make_map_update(**kw)
        w = {k: z["case%d_%s" % (n, k)] for k in ("kf_TcwGBA", "kf_visited", "mp_pos", "mp_corrected")}
        _same_as(oracle.gba_map_update(sc), w)
        _same_as(oracle.gba_map_update(sc, fn=host), w)
        assert w["mp_corrected"].sum() > 100 and (w["kf_visited"] == 1).sum() > 20


@pytest.mark.parametrize("kw", [dict(K=150, P=2500, seed=31), dict(K=300, P=1000, seed=32, chain=1.0, n_origins=1, new_kf_frac=0.5),
                                dict(K=25, P=3000, seed=33, new_kf_frac=0.6, n_origins=3)])
def test_oracle_against_live_opencv(oracle, kw):
    """the same witness on fresh scenes, where cv2 can be imported (it cannot be assumed on every box: the fixture above travels)"""
    pytest.importorskip("cv2")
    mod = _golden_cases()
    sc = synth.make_map_update(**kw)
    _same_as(oracle.gba_map_update(sc), mod.cv2_update(sc))


def test_gemm_rounding_is_f32_left_to_right():
    """the one fact the restatement rests on, checked directly: cv::gemm with inner dimension 3 or 4 sums f32 products left to right in f32"""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(8)
    for shape in ((4, 4, 4), (3, 3, 1), (3, 3, 3), (4, 4, 1)):
        for _ in range(300):
            A = (rng.normal(0, 1, shape[:2]) * 10.0 ** rng.integers(-3, 4)).astype(np.float32)
            B = (rng.normal(0, 1, (shape[1], shape[2])) * 10.0 ** rng.integers(-3, 4)).astype(np.float32)
            want = np.zeros((shape[0], shape[2]), np.float32)
            for i in range(shape[0]):
                for j in range(shape[2]):
                    s = np.float32(A[i, 0] * B[0, j])
                    for k in range(1, shape[1]):
                        s = np.float32(s + np.float32(A[i, k] * B[k, j]))
                    want[i, j] = s
            assert np.array_equal(cv2.gemm(A, B, 1.0, None, 0.0), want)
