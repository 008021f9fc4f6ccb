"""CPU suite: the arithmetic of the on-device window search (CCM_MATCH_WINDOW=1, k_window_best) run on the host.
tests/host/window_best_host.cpp compiles the product's csrc/window_best.cuh with g++ and runs the kernel's 32 lanes one after
another; Fuse x2 and SearchBySim3 built from those winners must equal the oracle index for index — including the tie-storm
scenes where only the visiting position packed into the key decides.  What this cannot show is the launch itself (uploads,
grid size, the shuffle instructions): that is tools/validate_prepared.sh step 1c on a device."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from ccm_slam_b200 import synth_match as sm
from ccm_slam_b200.frontend import grid_struct, queries_struct

HERE = os.path.dirname(os.path.abspath(__file__))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def wb(tmp_path_factory):
    cuda_inc = next((d for d in ("/usr/local/cuda/include", "/usr/local/cuda/targets/x86_64-linux/include") if os.path.exists(os.path.join(d, "vector_types.h"))), None)
    if cuda_inc is None:
        pytest.skip("vector_types.h (CUDA toolkit headers) not found")
    so = str(tmp_path_factory.mktemp("wb") / "libwindow_best_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-fno-fast-math", "-ffp-contract=off", "-I", cuda_inc,
                           "-o", so, os.path.join(HERE, "host", "window_best_host.cpp")])
    return C.CDLL(so)


def fuse(wb, g, q, w):
    keep = []; G = grid_struct(g, keep); Q = queries_struct(q, keep)
    w = None if w is None else np.ascontiguousarray(w, np.float32)
    best = np.empty(Q.m, np.int32); n = C.c_int32()
    assert wb.wb_fuse(C.byref(G), C.byref(Q), _p(w), 0 if w is None else len(w), _p(best), C.byref(n)) == 0
    return best, n.value


def by_sim3(wb, g1, g2, q12, q21):
    keep = []; G1 = grid_struct(g1, keep); G2 = grid_struct(g2, keep); Q12 = queries_struct(q12, keep); Q21 = queries_struct(q21, keep)
    out = np.empty(Q12.m, np.int32); n = C.c_int32()
    assert wb.wb_by_sim3(C.byref(G1), C.byref(G2), C.byref(Q12), C.byref(Q21), _p(out), C.byref(n)) == 0
    return out, n.value


@pytest.mark.parametrize("seed,n,m,th,ties", [(0, 1000, 1500, 3.0, False), (1, 2000, 3000, 7.0, False), (2, 300, 200, 15.0, False),
                                              (3, 1200, 1800, 6.0, True), (4, 2500, 2500, 12.0, True), (5, 40, 90, 30.0, True)])
def test_fuse_from_lane_walk(oracle, wb, seed, n, m, th, ties):
    g = sm.make_grid(n=n, seed=10 + seed, clustered=seed != 2)
    q = sm.make_queries(g, m=m, seed=20 + seed, th=th)
    if ties:
        g, q = sm.tie_storm(g, q, pool=6 if seed != 5 else 2, seed=40 + seed)
    for w in (None, sm.INV_LEVEL_SIGMA2):
        best, nf = fuse(wb, g, q, w)
        rbest, rn = oracle.fuse_search(g, q, w)
        assert nf == rn and np.array_equal(best, rbest)
        assert nf > (5 if n < 100 else 20)


def test_window_winners_carry_distances(oracle, wb):
    """index and distance of every window's winner against a brute-force scan of the distance matrix in visiting order"""
    g = sm.make_grid(n=800, seed=3); q = sm.make_queries(g, m=600, seed=4, th=9.0)
    g, q = sm.tie_storm(g, q, pool=4, seed=5)
    keep = []; G = grid_struct(g, keep); Q = queries_struct(q, keep)
    bi = np.empty(Q.m, np.int32); bd = np.empty(Q.m, np.int32)
    assert wb.wb_windows(C.byref(G), C.byref(Q), None, 0, _p(bi), _p(bd)) == 0
    D = np.unpackbits(np.asarray(q["desc"])[:, None, :] ^ np.asarray(g["desc"])[None, :, :], axis=2).sum(axis=2)
    from ccm_slam_b200.frontend import GetFeaturesInArea
    seen = 0
    for i in range(Q.m):
        if not q["valid"][i]:
            assert bi[i] == -1
            continue
        L = int(q["level"][i])
        cand = GetFeaturesInArea(g, float(q["uv"][i, 0]), float(q["uv"][i, 1]), float(q["radius"][i]), L - 1, L)
        if len(cand) == 0:
            assert bi[i] == -1 and bd[i] == 0x7fffffff
            continue
        d = D[i, cand]
        k = int(np.argmin(d))           # first minimum in visiting order
        assert bi[i] == cand[k] and bd[i] == d[k]
        seen += (d == d[k]).sum() > 1
    assert seen > 50                    # ties really occurred


@pytest.mark.parametrize("ties", [False, True])
def test_by_sim3_from_lane_walk(oracle, wb, ties):
    rng = np.random.default_rng(9)
    g1 = sm.make_grid(n=900, seed=7); g2 = sm.make_grid(n=950, seed=8)
    share = rng.permutation(900)[:500]
    g2["desc"][:500] = sm.flip_bits(g1["desc"][share], rng.integers(0, 30, 500), rng)
    g2["kp_xy"][:500] = g1["kp_xy"][share] + rng.normal(0, 1.5, (500, 2)).astype(np.float32)
    g2["octave"][:500] = g1["octave"][share]
    if ties:                            # few distinct descriptors on both sides: every window holds equal distances
        pool = rng.integers(0, 256, (5, 32)).astype(np.uint8)
        g1["desc"][:] = pool[rng.integers(0, 5, 900)]; g2["desc"][:] = pool[rng.integers(0, 5, 950)]

    def queries(src_g, dst_g, ps, pd):
        m = src_g["desc"].shape[0]
        uv = rng.uniform(0, 700, (m, 2)).astype(np.float32); level = src_g["octave"].copy()
        uv[ps] = dst_g["kp_xy"][pd] + rng.normal(0, 1.0, (len(ps), 2)).astype(np.float32)
        return dict(valid=(rng.random(m) < 0.8).astype(np.uint8), uv=uv, radius=(np.float32(7.5) * sm.SCALE_FACTORS[level]).astype(np.float32),
                    level=level, desc=src_g["desc"])
    q12 = queries(g1, g2, share, np.arange(500)); q21 = queries(g2, g1, np.arange(500), share)
    got, nf = by_sim3(wb, g1, g2, q12, q21)
    ref, rn = oracle.search_by_sim3(g1, g2, q12, q21)
    assert nf == rn and np.array_equal(got, ref) and nf > (20 if ties else 150)


def test_empty_and_out_of_image(oracle, wb):
    g = sm.make_grid(n=300, seed=1); q = sm.make_queries(g, m=50, seed=2)
    q["uv"][:10] = np.float32(-500.0); q["uv"][10:20] = np.float32(5000.0); q["valid"][20:25] = 0
    best, nf = fuse(wb, g, q, None)
    rbest, rn = oracle.fuse_search(g, q, None)
    assert nf == rn and np.array_equal(best, rbest) and (best[:25] == -1).all()
    e = {k: (v[:0] if isinstance(v, np.ndarray) else v) for k, v in q.items()}
    best, nf = fuse(wb, g, e, None)
    assert nf == 0 and len(best) == 0
