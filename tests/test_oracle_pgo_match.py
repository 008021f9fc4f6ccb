"""Pins the Sim3 / pose-graph and matcher parts of the CPU oracle with independent witnesses (scipy matrix exponential,
pure-Python restatements of the greedy matchers)."""
import numpy as np
import pytest
from scipy.linalg import expm
from scipy.spatial.transform import Rotation

from ccm_slam_b200 import synth


def _sim3_matrix(s8):
    M = np.eye(4)
    M[:3, :3] = s8[7] * Rotation.from_quat(s8[:4]).as_matrix()
    M[:3, 3] = s8[4:7]
    return M


def test_sim3_exp_is_the_matrix_exponential(oracle):
    rng = np.random.default_rng(0)
    for _ in range(40):
        u = np.concatenate([rng.normal(size=3) * rng.choice([1e-7, 0.3, 1.5]), rng.normal(size=3), [rng.normal() * rng.choice([1e-7, 0.2])]])
        G = np.zeros((4, 4))
        G[:3, :3] = np.array([[0, -u[2], u[1]], [u[2], 0, -u[0]], [-u[1], u[0], 0]]) + u[6] * np.eye(3)
        G[:3, 3] = u[3:6]
        th = np.linalg.norm(u[:3])
        got = _sim3_matrix(oracle.sim3_exp(u))
        # g2o's theta<1e-5 branch uses R = I + W + W^2, a 0.5*theta^2 deviation: far below 1e-9 here
        assert np.allclose(got, expm(G), atol=1e-9 if th < 1e-5 else 1e-12)
        assert np.allclose(oracle.sim3_log(oracle.sim3_exp(u)), u, atol=1e-9)


def test_sim3_group_ops_and_edge_error(oracle):
    rng = np.random.default_rng(1)
    mk = lambda: np.concatenate([Rotation.random(random_state=rng.integers(1 << 30)).as_quat(), rng.normal(size=3), [np.exp(rng.normal() * 0.1)]])
    for _ in range(20):
        a, b = mk(), mk()
        assert np.allclose(_sim3_matrix(oracle.sim3_mul(a, b)), _sim3_matrix(a) @ _sim3_matrix(b), atol=1e-12)
        assert np.allclose(_sim3_matrix(oracle.sim3_inv(a)), np.linalg.inv(_sim3_matrix(a)), atol=1e-12)
        meas = oracle.sim3_mul(b, oracle.sim3_inv(a))        # Sji = Sjw * Swi
        assert np.allclose(oracle.pgo_edge_error(meas, a, b), 0, atol=1e-12)


@pytest.mark.parametrize("fix_scale", [False, True])
def test_pose_graph_distributes_the_loop_error(oracle, fix_scale):
    p = synth.make_pgo(K=120, fix_scale=fix_scale)
    r = oracle.pgo_solve(p, iterations=20)
    # pose graphs are almost linear: one Gauss-Newton-like step (lambda0 = 1e-16) removes the distributable error, the
    # remaining chi2 is the inconsistency between the loop edges and the drifted odometry
    assert r["iters_done"] >= 1 and r["chi2_final"] < 0.1 * r["chi2_initial"]
    assert np.array_equal(r["sim3"][0], p.sim3[0])                       # the fixed loop keyframe
    if fix_scale:
        assert np.allclose(r["sim3"][:, 7], 1.0, atol=1e-12)
    assert np.isfinite(r["sim3"]).all()


def _py_bow_kf_frame(D, has, ak, fvk, af, fvf, nnratio, ori):
    """Independent restatement of SearchByBoW(KF, Frame) over a distance matrix (S/ORBmatcher.cpp:178-306)."""
    nf = D.shape[1]
    m = -np.ones(nf, int); hist = [[] for _ in range(30)]; n = 0
    nodes_f = {int(v): i for i, v in enumerate(fvf.node_id)}
    for a, nid in enumerate(fvk.node_id):
        if int(nid) not in nodes_f: continue
        b = nodes_f[int(nid)]
        for i in fvk.feat[fvk.node_ptr[a]:fvk.node_ptr[a + 1]]:
            if not has[i]: continue
            cand = [int(j) for j in fvf.feat[fvf.node_ptr[b]:fvf.node_ptr[b + 1]] if m[j] < 0]
            b1 = b2 = 256; bj = -1
            for j in cand:
                d = int(D[i, j])
                if d < b1: b2, b1, bj = b1, d, j
                elif d < b2: b2 = d
            if b1 <= 50 and np.float32(b1) < np.float32(nnratio) * np.float32(b2):
                m[bj] = i; n += 1
                if ori:
                    rot = np.float32(ak[i]) - np.float32(af[bj])
                    if rot < 0: rot += np.float32(360.0)
                    bn = int(np.floor(float(rot * np.float32(1.0 / 30)) + 0.5))
                    hist[0 if bn == 30 else bn].append(bj)
    if ori:
        sizes = [len(h) for h in hist]
        order = []
        m1 = m2 = m3 = 0; i1 = i2 = i3 = -1
        for i, s in enumerate(sizes):
            if s > m1: m3, m2, m1, i3, i2, i1 = m2, m1, s, i2, i1, i
            elif s > m2: m3, m2, i3, i2 = m2, s, i2, i
            elif s > m3: m3, i3 = s, i
        if m2 < 0.1 * m1: i2 = i3 = -1
        elif m3 < 0.1 * m1: i3 = -1
        for i, h in enumerate(hist):
            if i in (i1, i2, i3): continue
            for j in h: m[j] = -1; n -= 1
    return m, n


def test_search_by_bow_against_python_restatement(oracle):
    rng = np.random.default_rng(3)
    n1, n2 = 300, 320
    d1 = rng.integers(0, 256, size=(n1, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, size=(n2, 32), dtype=np.uint8)
    perm = rng.permutation(n2)[:200]
    d2[perm] = d1[:200]                                  # true matches ...
    flip = rng.integers(0, 32, size=(200, 3))
    for r, cols in zip(perm, flip):                      # ... with a few flipped bits
        d2[r, cols] ^= np.uint8(1 << int(rng.integers(0, 8)))
    node1 = rng.integers(0, 12, size=n1); node2 = rng.integers(0, 12, size=n2); node2[perm] = node1[:200]
    has = (rng.random(n1) < 0.8).astype(np.uint8)
    a1 = rng.uniform(0, 360, n1).astype(np.float32); a2 = rng.uniform(0, 360, n2).astype(np.float32); a2[perm] = a1[:200] + 2.0
    fv1, fv2 = oracle.FeatureVector(node1), oracle.FeatureVector(node2)
    D = np.unpackbits(d1[:, None, :] ^ d2[None, :, :], axis=2).sum(2)
    assert all(oracle.descriptor_distance(d1[i], d2[j]) == D[i, j] for i, j in [(0, 0), (5, 9), (299, 319)])
    assert oracle.descriptor_distance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256
    for nnratio, ori in [(0.7, True), (0.6, False)]:
        got, n = oracle.match_bow_kf_frame(d1, has, a1, fv1, d2, a2, fv2, nnratio, ori)
        ref, rn = _py_bow_kf_frame(D, has, a1, fv1, a2, fv2, nnratio, ori)
        assert n == rn and np.array_equal(got, ref) and n > 50
