"""CPU suite: shim/Optimizer_shim.cpp as a drop-in for cslam::Optimizer.

The shim is compiled against the reference's OWN cslam/Optimizer.h, Converter.h / Converter.cc, Datatypes.h, estd.h, config.h and g2o
value types (over the Eigen stand-in), with Map / KeyFrame / MapPoint / Frame replaced by stand-ins that carry the members the optimiser
code touches (oracle/ref_stub_opt), and driven through the class interface exactly as the reference's callers drive it
(oracle/ref_optimizer_wrap.cpp).  The device entry points are doubled by the CPU oracle (oracle/ccm_device_double.cpp), so what is under
test is everything the shim itself does: which keyframes, points and observations it selects, how it flattens them, which vertex it
fixes, what it writes back and into which field, how many times.  The expected values come from an independent restatement of the
reference's selection rules in this file plus the oracle on the resulting flat problem — bit for bit, since both sides run the same
optimiser.  Skipped where the reference tree or the product library is absent."""
import numpy as np
import pytest

from ccm_slam_b200 import api, synth
from tests import shim_optimizer_harness as H


@pytest.fixture(scope="module")
def ho(oracle):
    if H.lib() is None:
        pytest.skip("oracle/_ref/liboptimizer_shim.so not available (needs the reference tree and the product library)")
    return oracle


def gba_rule(sc):
    """MapFusionGBA's selection (S/Optimizer.cpp:695-787): non-bad keyframes; a point needs >= 2 observations and >= 2 of them on
    non-bad keyframes with mUniqueId <= the largest id among the keyframes added"""
    rows = [k for k in range(len(sc["kf_uid"])) if not sc["kf_bad"][k]]
    max_id = max(sc["kf_uid"][k] for k in rows)

    def rule(j, obs):
        if sc["mp_bad"][j] or len(obs) < 2:
            return None
        ok = [q for q in obs if not sc["kf_bad"][sc["obs_kf"][q]] and sc["kf_uid"][sc["obs_kf"][q]] <= max_id]
        return ok if len(ok) >= 2 else None
    return rows, rule


@pytest.mark.parametrize("loop_mode", [False, True])
@pytest.mark.parametrize("name,bad_kf,bad_mp", [("tiny", 0.0, 0.0), ("small", 0.0, 0.0), ("small", 0.1, 0.1)])
def test_map_fusion_gba(ho, name, bad_kf, bad_mp, loop_mode):
    p = synth.make_config(name)
    sc = H.scene_from_problem(p, ho, seed=3, map_id=0, bad_kf=bad_kf, bad_mp=bad_mp)
    sc["kf_bad"][0] = 0                                            # the origin keyframe stays
    rows, rule = gba_rule(sc)
    flat, mp_of_row = H.flat_from_scene(sc, ho, rows, [1 if k == sc["origin"] else 0 for k in rows], rule)
    want = ho.ba_solve(flat, iterations=8, robust=True, huber_delta=api.HUBER_GBA)
    loop = (7, 0) if loop_mode else (0, 0)                         # nLoopKF == (0, mMapId) means "write the poses directly"
    got = H.run_gba(sc, 0, 8, True, loop)
    K, P = len(sc["kf_uid"]), len(sc["mp_uid"])
    Tw = np.stack([ho.pose_to_Tcw_f32(q) for q in want["poses"]])
    Xw = want["points"].astype(np.float32)
    inc_kf = np.zeros(K, bool); inc_kf[rows] = True
    inc_mp = np.zeros(P, bool); inc_mp[mp_of_row] = True
    assert 0 < inc_mp.sum() <= P and (bad_mp == 0 or inc_mp.sum() < P)
    if not loop_mode:
        assert np.array_equal(got["kf_Tcw"][rows], Tw) and np.array_equal(got["kf_set_pose"], inc_kf.astype(np.int32))
        assert np.array_equal(got["kf_Tcw"][~inc_kf], sc["kf_Tcw"][~inc_kf])
        assert np.array_equal(got["mp_pos"][mp_of_row], Xw) and np.array_equal(got["mp_pos"][~inc_mp], sc["mp_pos"][~inc_mp])
        assert np.array_equal(got["mp_set_pos"], inc_mp.astype(np.int32)) and np.array_equal(got["mp_update_normal"], inc_mp.astype(np.int32))
        assert not got["kf_gba_tag"].any() and not got["mp_gba_tag"].any()
    else:
        assert np.array_equal(got["kf_TcwGBA"][rows], Tw) and np.array_equal(got["kf_Tcw"], sc["kf_Tcw"]) and not got["kf_set_pose"].any()
        assert np.array_equal(got["mp_posGBA"][mp_of_row], Xw) and np.array_equal(got["mp_pos"], sc["mp_pos"]) and not got["mp_set_pos"].any()
        assert np.array_equal(got["kf_gba_tag"][rows], np.tile(loop, (len(rows), 1))) and not got["kf_gba_tag"][~inc_kf].any()
        assert np.array_equal(got["mp_gba_tag"][mp_of_row], np.tile(loop, (len(mp_of_row), 1))) and not got["mp_gba_tag"][~inc_mp].any()
    assert np.abs(Tw - sc["kf_Tcw"][rows]).max() > 1e-4           # the optimisation did move things


@pytest.mark.parametrize("which,loop_mode", [(0, False), (0, True), (1, False)])
def test_gba_keyframe_turning_bad_during_the_solve(ho, which, loop_mode):
    """LoopFinder::RunGBA and the client GBA run in their own thread while culling continues: a keyframe that turns bad between flattening
    and write-back must not shift the rows of the keyframes after it (the reference looks every vertex up by id, S/Optimizer.cpp:805).
    Every other keyframe and every point gets exactly what it gets without the flip; the flipped keyframe is left alone."""
    p = synth.make_config("small")
    sc = H.scene_from_problem(p, ho, seed=3, map_id=0)
    loop = (7, 0) if loop_mode else (0, 0)
    base = H.run_gba(sc, which, 6, True, loop)
    K = len(sc["kf_uid"])
    flip = K // 2
    got = H.run_gba_flip(sc, which, 6, True, loop, flip)
    others = np.arange(K) != flip
    key = "kf_TcwGBA" if loop_mode else "kf_Tcw"
    assert np.array_equal(got[key][others], base[key][others])
    assert np.abs(base[key][flip + 1] - base[key][flip]).max() > 1e-3     # neighbouring rows do differ: a shift would have shown
    if loop_mode:
        assert not got["kf_TcwGBA"][flip].any() and not got["kf_gba_tag"][flip].any()
    else:
        assert np.array_equal(got["kf_Tcw"][flip], sc["kf_Tcw"][flip]) and got["kf_set_pose"][flip] == 0
    for k in ("mp_pos", "mp_posGBA", "mp_gba_tag", "mp_set_pos"):
        assert np.array_equal(got[k], base[k]), k


@pytest.mark.parametrize("loop_mode", [False, True])
@pytest.mark.parametrize("name,bad_kf,bad_mp", [("small", 0.0, 0.0), ("small", 0.1, 0.1)])
def test_map_fusion_gba_through_the_persistent_mirror(ho, name, bad_kf, bad_mp, loop_mode):
    """with a ccm_map_mirror registered for the map, MapFusionGBA takes the flat problem from the mirror (no walk
    over GetObservations()) and writes back by id.  Fed in map order the mirror hands out the very arrays the per-call flattening builds,
    so everything the call leaves in the map is bit-identical to the plain path."""
    p = synth.make_config(name)
    sc = H.scene_from_problem(p, ho, seed=5, map_id=0, bad_kf=bad_kf, bad_mp=bad_mp)
    sc["kf_bad"][0] = 0
    loop = (7, 0) if loop_mode else (0, 0)
    plain = H.run_gba(sc, 0, 8, True, loop)
    mirrored = H.run_gba_mirror(sc, 8, True, loop)
    for k in plain:
        assert np.array_equal(plain[k], mirrored[k]), k
    assert np.abs(plain["kf_TcwGBA" if loop_mode else "kf_Tcw"] - sc["kf_Tcw"]).max() > 1e-4


def test_second_gba_on_an_unchanged_structure_reuses_the_device_state(ho):
    """Two global BAs in a row on one map whose structure did not change in between (values only: the first BA's own results): with a
    registered mirror the second call keeps the solver handle -- structure, observations and everything built from them stay on the
    device, ccm_ba_set_estimate uploads the estimate -- and the map it leaves is bit-identical to the per-call path (which creates a
    handle inside ccm_ba_solve each time)."""
    p = synth.make_config("small")
    sc = H.scene_from_problem(p, ho, seed=9, map_id=0)
    plain, n_plain = H.run_gba_twice(sc, False, 4)
    mirrored, n_mirror = H.run_gba_twice(sc, True, 4)
    for k in plain:
        assert np.array_equal(plain[k], mirrored[k]), k
    assert n_mirror == 1                                           # one ccm_ba_create for two BAs (the plain path goes through ccm_ba_solve)
    assert (plain["kf_set_pose"] == 2).all()                       # both BAs wrote every keyframe


# ---- essential graph ----------------------------------------------------------------------------------------------------------
MIN_FEAT = 100      # params::opt::miEssGraphMinFeats, handed to the reference's config.h through the FileStorage stand-in (conftest sets it)


def essential_scene(ho, K=40, seed=0, bad=()):
    """a drifting chain of keyframes: spanning tree k -> k-1, covisibility with weights around the threshold, two earlier loop edges, a new
    loop connection between the last keyframe and an early one; points hang off reference keyframes"""
    rng = np.random.default_rng(seed)
    pg = synth.make_pgo(K=K, n_loop=0, seed=seed)
    # keyframe poses: the Sim3 rotations / translations of the synthetic graph as f32 Tcw (scale 1)
    Tcw = np.zeros((K, 4, 4), np.float32)
    for k in range(K):
        Tcw[k] = ho.pose_to_Tcw_f32(pg.sim3[k, :7])
    parent = np.arange(-1, K - 1).astype(np.int32)
    cov = [[] for _ in range(K)]
    for k in range(K):
        for j in range(max(0, k - 4), min(K, k + 5)):
            if j != k:
                cov[k].append((j, int(rng.choice([30, 99, 100, 150, 400]))))
        cov[k].sort(key=lambda t: -t[1])
    loops = {K - 10: [3], 3: [K - 10], K - 20: [1], 1: [K - 20]}     # mspLoopEdges is symmetric in the reference
    P = 300
    ref = rng.integers(0, K, P).astype(np.int32)
    pos = rng.normal(0, 3, (P, 3)).astype(np.float32)
    sc = dict(kf_uid=(np.arange(K) * 3 + 5).astype(np.int64), kf_id=np.stack([np.arange(K), np.zeros(K, np.int64)], 1), kf_bad=np.zeros(K, np.uint8),
              kf_Tcw=Tcw, kf_intr=np.tile(np.float32(synth.EUROC_INTR), (K, 1)), kp_ptr=np.zeros(K + 1, np.int32), kp_uv=np.zeros((0, 2), np.float32),
              kp_octave=np.zeros(0, np.int32), inv_level_sigma2=H.sm.INV_LEVEL_SIGMA2, kf_parent=parent,
              loop_ptr=np.concatenate([[0], np.cumsum([len(loops.get(k, [])) for k in range(K)])]).astype(np.int32),
              loop_kf=np.array([j for k in range(K) for j in loops.get(k, [])], np.int32),
              cov_ptr=np.concatenate([[0], np.cumsum([len(c) for c in cov])]).astype(np.int32),
              cov_kf=np.array([j for c in cov for j, _ in c], np.int32), cov_w=np.array([w for c in cov for _, w in c], np.int32),
              mp_uid=np.arange(P).astype(np.int64) + 9000000, mp_id=np.stack([np.arange(P), np.zeros(P, np.int64)], 1), mp_bad=(rng.random(P) < 0.05).astype(np.uint8),
              mp_pos=pos, mp_ref=ref, obs_ptr=np.zeros(P + 1, np.int32), obs_kf=np.zeros(0, np.int32), obs_idx=np.zeros(0, np.int32), origin=0, map_id=0)
    for k in bad:
        sc["kf_bad"][k] = 1
    return sc, cov, loops


def expected_essential_graph(ho, sc, cov, loops, loop_kf, cur_kf, conn, fix_scale, corrected=None, noncorrected=None, mp_corr_ref=None):
    """the reference's graph (S/Optimizer.cpp:1086-1268 / 1360-1504) on flat arrays, solved by the oracle, recovered as :1280-1330"""
    R = ho.Pieces("ref"); O = ho.Pieces("oracle")
    K = len(sc["kf_uid"]); uid = sc["kf_uid"]; bad = sc["kf_bad"].astype(bool)
    corrected = corrected or {}; noncorrected = noncorrected or {}
    Scw = {}
    for k in range(K):
        if bad[k]:
            continue
        T = sc["kf_Tcw"][k].astype(np.float64)
        Scw[k] = corrected[k] if k in corrected else R.vec("sim3_from_Rt", 8, T[:3, :3].ravel(), T[:3, 3], [1.0])
    rows = sorted(Scw, key=lambda k: uid[k]); row_of = {k: r for r, k in enumerate(rows)}
    unc = lambda k: noncorrected[k] if k in noncorrected else Scw[k]
    mul = lambda a, b: O.vec("sim3_mul", 8, a, b); inv = lambda a: O.vec("sim3_inv", 8, a)
    ei, ej, meas, inserted = [], [], [], set()

    def add(i, j, S):
        if i in row_of and j in row_of:
            ei.append(row_of[i]); ej.append(row_of[j]); meas.append(S)
    weight = lambda i, j: next((w for jj, w in cov[i] if jj == j), 0)
    for i in sorted(conn):
        if bad[i]:
            continue
        Swi = inv(Scw[i])
        for j in sorted(conn[i]):
            if bad[j]:
                continue
            if (uid[i] != uid[cur_kf] or uid[j] != uid[loop_kf]) and weight(i, j) < MIN_FEAT:
                continue
            add(i, j, mul(Scw[j], Swi)); inserted.add((min(uid[i], uid[j]), max(uid[i], uid[j])))
    children = {k: set() for k in range(K)}
    for k in range(K):
        if sc["kf_parent"][k] >= 0:
            children[int(sc["kf_parent"][k])].add(k)
    for i in range(K):
        if bad[i]:
            continue
        Swi = inv(unc(i)); par = int(sc["kf_parent"][i])
        if par >= 0:
            add(i, par, mul(unc(par), Swi) if par in Scw else None)
        le = sorted(loops.get(i, []))
        for j in le:
            if uid[j] < uid[i]:
                add(i, j, mul(unc(j), Swi) if j in Scw else None)
        for j, w in cov[i]:
            if w < MIN_FEAT or bad[j]:
                continue
            if j != par and j not in children[i] and j not in le and uid[j] < uid[i]:
                if (min(uid[i], uid[j]), max(uid[i], uid[j])) in inserted:
                    continue
                add(i, j, mul(unc(j), Swi))
    pg = synth.PGOProblem(sim3=np.stack([Scw[k] for k in rows]), fixed=np.array([1 if k == loop_kf else 0 for k in rows], np.uint8),
                          edge_i=np.array(ei, np.int32), edge_j=np.array(ej, np.int32), meas=np.stack(meas), fix_scale=bool(fix_scale))
    res = ho.pgo_solve(pg, iterations=20, lambda_init=1e-16)
    Tnew = {}; Swc = {}
    for r, k in enumerate(rows):
        S = res["sim3"][r]; Swc[k] = inv(S)
        M = R.vec("se3_homogeneous", 16, np.r_[S[:4], S[4:7]]).reshape(4, 4)      # rotation matrix of the quaternion; translation replaced below
        T = np.eye(4, dtype=np.float32); T[:3, :3] = M[:3, :3].astype(np.float32)
        t = S[4:7].copy(); t *= (1.0 / S[7]); T[:3, 3] = t.astype(np.float32)
        Tnew[k] = T
    Xnew = {}
    for j in range(len(sc["mp_uid"])):
        if sc["mp_bad"][j]:
            continue
        kr = int(sc["mp_ref"][j])
        if mp_corr_ref is not None and mp_corr_ref[j] >= 0:
            kr = int(np.flatnonzero(uid == mp_corr_ref[j])[0])
        if kr not in Scw:
            continue
        Xnew[j] = O.vec("sim3_map", 3, Swc[kr], O.vec("sim3_map", 3, Scw[kr], sc["mp_pos"][j].astype(np.float64))).astype(np.float32)
    return pg, res, Tnew, Xnew


@pytest.mark.parametrize("fix_scale", [False, True])
@pytest.mark.parametrize("bad", [(), (7, 22)])
def test_essential_graph_map_fusion(ho, fix_scale, bad):
    K = 40
    sc, cov, loops = essential_scene(ho, K=K, seed=1, bad=bad)
    loop_kf, cur_kf = 2, K - 1
    conn = {cur_kf: [loop_kf, 4], loop_kf: [cur_kf], 4: [cur_kf], K - 2: [3]}     # the pair (cur, loop) passes regardless of its weight
    pg, res, Tnew, Xnew = expected_essential_graph(ho, sc, cov, loops, loop_kf, cur_kf, conn, fix_scale)
    got = H.run_essential_graph(sc, loop_kf, cur_kf, conn, fix_scale)
    assert pg.edge_i.size > 2 * K and res["iters_done"] >= 1
    for k in range(K):
        if k in Tnew:
            assert np.array_equal(got["kf_Tcw"][k], Tnew[k]), k
            assert got["kf_set_pose"][k] == 1
        else:
            assert got["kf_set_pose"][k] == 0 and np.array_equal(got["kf_Tcw"][k], sc["kf_Tcw"][k])
    for j in range(len(sc["mp_uid"])):
        if j in Xnew:
            assert np.array_equal(got["mp_pos"][j], Xnew[j]) and got["mp_set_pos"][j] == 1 and got["mp_update_normal"][j] == 1
        else:
            assert got["mp_set_pos"][j] == 0
    # every measurement of this variant is composed from the current poses: the graph is consistent and the solve leaves it where it was
    # (up to the f32 round trip) — as in the reference; the loop-closure variant below is the one that moves things
    assert res["chi2_initial"] < 1e-9


def test_essential_graph_loop_closure(ho):
    """the loop-closure variant: keyframes around the current one come with corrected / non-corrected Sim3s (CorrectLoop's maps); points
    corrected through the current keyframe name their reference explicitly"""
    K = 40
    sc, cov, loops = essential_scene(ho, K=K, seed=2)
    loop_kf, cur_kf = 2, K - 1
    conn = {cur_kf: [loop_kf, 4], loop_kf: [cur_kf], 4: [cur_kf]}
    O = ho.Pieces("oracle"); R = ho.Pieces("ref")
    rng = np.random.default_rng(5)
    near = [K - 1, K - 2, K - 3]
    fix = O.vec("sim3_exp", 8, np.r_[rng.normal(0, 0.02, 3), rng.normal(0, 0.1, 3), 0.05])
    non = {}; cor = {}
    for k in near:
        T = sc["kf_Tcw"][k].astype(np.float64)
        non[k] = R.vec("sim3_from_Rt", 8, T[:3, :3].ravel(), T[:3, 3], [1.0])
        cor[k] = O.vec("sim3_mul", 8, non[k], fix)
    mp_corr = np.where(np.random.default_rng(6).random(300) < 0.2, sc["kf_uid"][K - 2], -1).astype(np.int32)
    pg, res, Tnew, Xnew = expected_essential_graph(ho, sc, cov, loops, loop_kf, cur_kf, conn, False, corrected=cor, noncorrected=non, mp_corr_ref=mp_corr)
    got = H.run_essential_graph(sc, loop_kf, cur_kf, conn, False, loop_closure=True,
                                corr=(near, np.stack([cor[k] for k in near]), np.stack([non[k] for k in near])), mp_corr_ref=mp_corr)
    for k in range(K):
        assert np.array_equal(got["kf_Tcw"][k], Tnew[k]), k
    for j, X in Xnew.items():
        assert np.array_equal(got["mp_pos"][j], X), j
    assert (mp_corr >= 0).sum() > 20 and res["chi2_initial"] > 1e-6 and res["chi2_final"] < res["chi2_initial"]
    assert max(np.abs(Tnew[k] - sc["kf_Tcw"][k]).max() for k in Tnew) > 1e-3


# ---- the BA client, local BA and the two single-vertex entry points -------------------------------------------------------------
def test_global_bundle_adjustment_client(ho):
    """BundleAdjustmentClient (S/Optimizer.cpp:40-212): every non-bad keyframe, fixed = keyframe (0, ClientId); a point needs one valid observation"""
    p = synth.make_config("small")
    sc = H.scene_from_problem(p, ho, seed=4, map_id=0, bad_kf=0.1, bad_mp=0.1)
    sc["kf_bad"][0] = 0
    rows = [k for k in range(p.K) if not sc["kf_bad"][k]]

    def rule(j, obs):
        if sc["mp_bad"][j]:
            return None
        ok = [q for q in obs if not sc["kf_bad"][sc["obs_kf"][q]]]
        return ok if ok else None
    flat, mp_of_row = H.flat_from_scene(sc, ho, rows, [1 if tuple(sc["kf_id"][k]) == (0, 0) else 0 for k in rows], rule)
    want = ho.ba_solve(flat, iterations=5, robust=True, huber_delta=api.HUBER_GBA)
    got = H.run_gba(sc, 1, 5, True, (0, 0))
    assert np.array_equal(got["kf_Tcw"][rows], np.stack([ho.pose_to_Tcw_f32(q) for q in want["poses"]]))
    assert np.array_equal(got["mp_pos"][mp_of_row], want["points"].astype(np.float32))
    inc = np.zeros(p.P, np.int32); inc[mp_of_row] = 1
    assert np.array_equal(got["mp_update_normal"], inc) and 0 < inc.sum() < p.P


def test_local_bundle_adjustment_client(ho):
    """LocalBundleAdjustmentClient (S/Optimizer.cpp:349-644): window selection, optimize(5), chi2 / depth classification, kernels dropped,
    optimize(10) from where round 1 stopped, observations erased, poses and points written back"""
    p = synth.make_config("cfg2", P=500)
    sc = H.scene_from_problem(p, ho, seed=5, map_id=0)
    K, P = p.K, p.P
    rng = np.random.default_rng(6)
    center = 3
    covis = [k for k in rng.permutation(K) if k != center][:10]
    sc["kf_bad"][covis[2]] = 1                                      # a bad covisible keyframe is tagged but stays out
    cov_ptr = np.zeros(K + 1, np.int32); cov_ptr[center + 1:] = len(covis)
    sc.update(cov_ptr=cov_ptr, cov_kf=np.array(covis, np.int32), cov_w=np.full(len(covis), 200, np.int32))
    sc["mp_bad"] = (rng.random(P) < 0.05).astype(np.uint8)
    # the window, restated
    local = [center] + [k for k in covis if not sc["kf_bad"][k]]
    tagged_local = set([center] + covis)
    kp_of = lambda k: range(sc["kp_ptr"][k], sc["kp_ptr"][k + 1])
    mp_at = {}                                                      # (keyframe, keypoint index) -> point
    for j in range(P):
        for q in range(sc["obs_ptr"][j], sc["obs_ptr"][j + 1]):
            mp_at[(int(sc["obs_kf"][q]), int(sc["obs_idx"][q]))] = j
    local_mps, seen = [], set()
    for k in local:
        for i in range(sc["kp_ptr"][k + 1] - sc["kp_ptr"][k]):
            j = mp_at.get((k, i))
            if j is not None and not sc["mp_bad"][j] and j not in seen:
                local_mps.append(j); seen.add(j)
    fixed_cams, tagged_fixed = [], set()
    for j in local_mps:
        for q in range(sc["obs_ptr"][j], sc["obs_ptr"][j + 1]):
            k = int(sc["obs_kf"][q])
            if k not in tagged_local and k not in tagged_fixed:
                tagged_fixed.add(k)
                if not sc["kf_bad"][k]:
                    fixed_cams.append(k)
    rows = local + fixed_cams
    fixed = [1 if tuple(sc["kf_id"][k]) == (0, 0) else 0 for k in local] + [1] * len(fixed_cams)
    in_rows = set(rows)
    order = {j: r for r, j in enumerate(local_mps)}

    def rule(j, obs):
        return [q for q in obs if not sc["kf_bad"][sc["obs_kf"][q]] and int(sc["obs_kf"][q]) in in_rows] if j in order else None
    # flat_from_scene walks points in index order; the shim walks lLocalMapPoints — build rows in that order instead
    sc_perm = dict(sc)
    flat, mp_of_row = H.flat_from_scene(sc, ho, rows, fixed, rule)
    perm = np.argsort([order[j] for j in mp_of_row])               # row r of the shim = local_mps[r]
    inv = np.empty_like(perm); inv[perm] = np.arange(len(perm))
    e_order = np.lexsort((np.arange(flat.E), inv[flat.obs_mp]))
    flat2 = synth.BAProblem(poses=flat.poses, intr=flat.intr, fixed=flat.fixed, points=flat.points[perm], obs_kf=flat.obs_kf[e_order],
                            obs_mp=inv[flat.obs_mp][e_order].astype(np.int32), obs_uv=flat.obs_uv[e_order], obs_w=flat.obs_w[e_order])
    r1 = ho.ba_solve(flat2, iterations=5, robust=True, huber_delta=api.HUBER_LOCAL)
    out = (r1["chi2"] > 5.991) | (r1["depth_pos"] == 0)
    p2 = flat2.copy(); p2.poses = r1["poses"]; p2.points = r1["points"]; p2.edge_flags = (out.astype(np.uint8) | 2).astype(np.uint8)
    r2 = ho.ba_solve(p2, iterations=10, robust=True, huber_delta=api.HUBER_LOCAL, chi2_in=r1["chi2"])
    erase = (r2["chi2"] > 5.991) | (r2["depth_pos"] == 0)
    got = H.run_local_ba(sc, center)
    assert np.array_equal(got["kf_Tcw"][local], np.stack([ho.pose_to_Tcw_f32(q) for q in r2["poses"][:len(local)]]))
    untouched = [k for k in range(K) if k not in local]
    assert np.array_equal(got["kf_Tcw"][untouched], sc["kf_Tcw"][untouched]) and not got["kf_set_pose"][untouched].any()
    assert np.array_equal(got["mp_pos"][local_mps], r2["points"].astype(np.float32))
    assert np.array_equal(got["mp_update_normal"][local_mps], np.ones(len(local_mps), np.int32))
    n_obs0 = np.diff(sc["obs_ptr"])
    erased_per_mp = np.bincount(np.asarray(local_mps)[flat2.obs_mp[erase]], minlength=P)
    assert np.array_equal(n_obs0 - got["mp_n_obs"], erased_per_mp) and erase.sum() > 0 and len(fixed_cams) > 0 and out.sum() > 0


@pytest.mark.parametrize("n,seed,frac", [(300, 11, 0.15), (40, 12, 0.3), (2, 13, 0.0)])
def test_pose_optimization_client(ho, n, seed, frac):
    d = synth.make_pose_opt(n=n, seed=seed, outlier_frac=frac)
    rng = np.random.default_rng(seed)
    extra = 25                                                      # keypoints without a map point: skipped, flags left alone
    octave = rng.integers(0, 8, n).astype(np.int32)
    w = H.sm.INV_LEVEL_SIGMA2[octave]
    T32 = ho.pose_to_Tcw_f32(d["Tcw0"])
    sc = dict(kf_uid=np.zeros(1, np.int64), kf_id=np.zeros((1, 2), np.int64), kf_bad=np.zeros(1, np.uint8), kf_Tcw=np.eye(4, dtype=np.float32)[None],
              kf_intr=np.float32(d["intr"])[None], kp_ptr=np.zeros(2, np.int32), kp_uv=np.zeros((0, 2), np.float32), kp_octave=np.zeros(0, np.int32),
              inv_level_sigma2=H.sm.INV_LEVEL_SIGMA2, kf_parent=None, loop_ptr=None, loop_kf=None, cov_ptr=None, cov_kf=None, cov_w=None,
              mp_uid=np.arange(n).astype(np.int64), mp_id=np.stack([np.arange(n), np.zeros(n, np.int64)], 1), mp_bad=np.zeros(n, np.uint8),
              mp_pos=np.float32(d["Xw"]), mp_ref=np.zeros(n, np.int32), obs_ptr=np.zeros(n + 1, np.int32), obs_kf=np.zeros(0, np.int32),
              obs_idx=np.zeros(0, np.int32), origin=0, map_id=0)
    slot = rng.permutation(n + extra)[:n]                           # where in the frame's keypoint list correspondence e sits
    slot.sort()
    kp_uv = rng.uniform(0, 700, (n + extra, 2)).astype(np.float32); kp_oct = rng.integers(0, 8, n + extra).astype(np.int32)
    mp_of_kp = np.full(n + extra, -1, np.int32)
    kp_uv[slot] = d["uv"]; kp_oct[slot] = octave; mp_of_kp[slot] = np.arange(n)
    r, Tout, outl, nsp = H.run_pose_optimization(sc, kp_uv, kp_oct, mp_of_kp, T32, d["intr"])
    Tq = ho.pose_from_Tcw_f32(T32)
    wT, wout, wn = ho.pose_optimize(Tq, np.float32(d["Xw"]), np.float32(d["uv"]), w, d["intr"])
    assert r == wn
    rest = np.setdiff1d(np.arange(n + extra), slot)
    if n < 3:
        assert r == 0 and nsp == 0 and np.array_equal(Tout, T32)
        assert not outl[slot].any()                                 # reset before the early return (S/Optimizer.cpp:262)
    else:
        assert nsp == 1 and np.array_equal(Tout, ho.pose_to_Tcw_f32(wT)) and np.array_equal(outl[slot], wout)
    assert outl[rest].all()                                         # never touched


@pytest.mark.parametrize("n,seed,fix,frac", [(120, 12, False, 0.2), (120, 12, True, 0.2), (12, 5, False, 0.6)])
def test_optimize_sim3(ho, n, seed, fix, frac):
    """OptimizeSim3: both keyframes at the identity pose, so the camera-frame points the shim forms (R*P + t in f32) are the stored positions"""
    d = synth.make_sim3_opt(n=n, seed=seed, fix_scale=fix, outlier_frac=frac)
    rng = np.random.default_rng(seed + 1)
    o1 = rng.integers(0, 8, n).astype(np.int32); o2 = rng.integers(0, 8, n).astype(np.int32)
    w1 = H.sm.INV_LEVEL_SIGMA2[o1]; w2 = H.sm.INV_LEVEL_SIGMA2[o2]
    P1, P2 = np.float32(d["P1c"]), np.float32(d["P2c"])
    I = np.eye(4, dtype=np.float32)
    match1 = (n + np.arange(n)).astype(np.int32)
    drop = rng.random(n) < 0.1; match1[drop] = -1                    # no match for this keypoint
    bad = np.zeros(2 * n, np.uint8); bad[rng.permutation(2 * n)[:n // 10]] = 1
    unseen = rng.random(n) < 0.05                                   # the matched point is not observed by keyframe 2: GetIndexInKeyFrame < 0
    obs_cnt = np.r_[np.ones(n, np.int32), (~unseen).astype(np.int32)]
    obs_ptr = np.concatenate([[0], np.cumsum(obs_cnt)]).astype(np.int32)
    obs_kf = np.r_[np.zeros(n, np.int32), np.ones((~unseen).sum(), np.int32)]
    obs_idx = np.r_[np.arange(n), np.arange(n)[~unseen]].astype(np.int32)
    intr = np.stack([np.float32(d["K1"]), np.float32(d["K2"])])
    sc = dict(kf_uid=np.arange(2).astype(np.int64), kf_id=np.stack([np.arange(2), np.zeros(2, np.int64)], 1), kf_bad=np.zeros(2, np.uint8),
              kf_Tcw=np.stack([I, I]), kf_intr=intr, kp_ptr=np.array([0, n, 2 * n], np.int32), kp_uv=np.r_[np.float32(d["uv1"]), np.float32(d["uv2"])],
              kp_octave=np.r_[o1, o2], inv_level_sigma2=H.sm.INV_LEVEL_SIGMA2, kf_parent=None, loop_ptr=None, loop_kf=None, cov_ptr=None,
              cov_kf=None, cov_w=None, mp_uid=np.arange(2 * n).astype(np.int64), mp_id=np.stack([np.arange(2 * n), np.zeros(2 * n, np.int64)], 1),
              mp_bad=bad, mp_pos=np.r_[P1, P2], mp_ref=np.zeros(2 * n, np.int32), obs_ptr=obs_ptr, obs_kf=obs_kf, obs_idx=obs_idx, origin=0, map_id=0)
    r, S, m_out = H.run_optimize_sim3(sc, 0, 1, match1, d["S12_0"], d["th2"], fix)
    sel = (match1 >= 0) & (bad[:n] == 0) & (bad[n:] == 0) & ~unseen      # the reference's pair filter (:917-931)
    wS, winl, wn = ho.sim3_optimize(d["S12_0"], P1[sel], P2[sel], np.float32(d["uv1"])[sel], np.float32(d["uv2"])[sel], w1[sel], w2[sel],
                                    d["K1"], d["K2"], d["th2"], fix)
    assert r == wn
    exp = match1.copy(); idx = np.flatnonzero(sel); exp[idx[winl == 0]] = -1
    assert np.array_equal(m_out, exp)
    assert np.array_equal(S, wS if wn > 0 or not np.array_equal(wS, d["S12_0"]) else np.asarray(d["S12_0"], np.float64))
    if fix and wn:
        assert S[7] == d["S12_0"][7]


# ---- against the reference's OWN Optimizer.cpp ----------------------------------------------------------------------------------
# oracle/_ref/liboptimizer_ref.so is cslam/src/Optimizer.cpp itself, compiled where it lies together with the whole of g2o's core and
# types from the reference tree (over the Eigen stand-in), the same stand-in Map / KeyFrame / MapPoint / Frame and the same wrapper; the
# only parts that are not the reference's are LinearSolverEigen / LinearSolverDense, backed by the oracle's factorisations.  Every entry
# point is driven through both libraries on the same scene and everything the callee wrote must be identical, bit for bit.
class BothSides:
    def __init__(self):
        self.n = 0

    def run(self, fn, *a, **k):
        H.use_reference(False); shim = fn(*a, **k)
        H.use_reference(True)
        try:
            ref = fn(*a, **k)
        finally:
            H.use_reference(False)
        self.n += 1
        return shim, ref


@pytest.fixture(scope="module")
def both(ho):
    H.use_reference(True)
    try:
        ok = H.lib() is not None
    finally:
        H.use_reference(False)
    if not ok:
        pytest.skip("oracle/_ref/liboptimizer_ref.so not available")
    return BothSides()


def same_out(a, b):
    if isinstance(a, dict):
        for k in a:
            assert np.array_equal(a[k], b[k]), k
    else:
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert np.array_equal(np.asarray(x), np.asarray(y))


@pytest.mark.parametrize("loop", [(0, 0), (7, 0)])
@pytest.mark.parametrize("name,bad_kf,bad_mp,robust", [("tiny", 0.0, 0.0, True), ("small", 0.0, 0.0, True), ("small", 0.1, 0.1, True), ("small", 0.1, 0.1, False)])
def test_reference_map_fusion_gba(ho, both, name, bad_kf, bad_mp, robust, loop):
    p = synth.make_config(name)
    sc = H.scene_from_problem(p, ho, seed=3, map_id=0, bad_kf=bad_kf, bad_mp=bad_mp)
    sc["kf_bad"][0] = 0
    for which in (0, 1):                                   # MapFusionGBA, GlobalBundleAdjustemntClient
        same_out(*both.run(H.run_gba, sc, which, 8, robust, loop))


def test_reference_local_ba(ho, both):
    p = synth.make_config("cfg2", P=500)
    sc = H.scene_from_problem(p, ho, seed=5, map_id=0)
    rng = np.random.default_rng(6)
    for center in (3, 11):
        covis = [k for k in rng.permutation(p.K) if k != center][:10]
        s2 = dict(sc); s2["kf_bad"] = sc["kf_bad"].copy(); s2["kf_bad"][covis[2]] = 1
        cov_ptr = np.zeros(p.K + 1, np.int32); cov_ptr[center + 1:] = len(covis)
        s2.update(cov_ptr=cov_ptr, cov_kf=np.array(covis, np.int32), cov_w=np.full(len(covis), 200, np.int32), mp_bad=(rng.random(p.P) < 0.05).astype(np.uint8))
        for server in (False, True):
            a, b = both.run(H.run_local_ba, s2, center, server)
            same_out(a, b)
            assert a["kf_set_pose"].sum() > 3 and (np.diff(s2["obs_ptr"]) - a["mp_n_obs"]).sum() > 0     # poses written, observations erased


@pytest.mark.parametrize("n,seed,frac", [(300, 11, 0.15), (40, 12, 0.3), (9, 14, 0.0), (2, 13, 0.0)])
def test_reference_pose_optimization(ho, both, n, seed, frac):
    d = synth.make_pose_opt(n=n, seed=seed, outlier_frac=frac)
    rng = np.random.default_rng(seed)
    extra = 25
    octave = rng.integers(0, 8, n).astype(np.int32)
    T32 = ho.pose_to_Tcw_f32(d["Tcw0"])
    sc = dict(kf_uid=np.zeros(1, np.int64), kf_id=np.zeros((1, 2), np.int64), kf_bad=np.zeros(1, np.uint8), kf_Tcw=np.eye(4, dtype=np.float32)[None],
              kf_intr=np.float32(d["intr"])[None], kp_ptr=np.zeros(2, np.int32), kp_uv=np.zeros((0, 2), np.float32), kp_octave=np.zeros(0, np.int32),
              inv_level_sigma2=H.sm.INV_LEVEL_SIGMA2, kf_parent=None, loop_ptr=None, loop_kf=None, cov_ptr=None, cov_kf=None, cov_w=None,
              mp_uid=np.arange(n).astype(np.int64), mp_id=np.stack([np.arange(n), np.zeros(n, np.int64)], 1), mp_bad=np.zeros(n, np.uint8),
              mp_pos=np.float32(d["Xw"]), mp_ref=np.zeros(n, np.int32), obs_ptr=np.zeros(n + 1, np.int32), obs_kf=np.zeros(0, np.int32),
              obs_idx=np.zeros(0, np.int32), origin=0, map_id=0)
    slot = np.sort(rng.permutation(n + extra)[:n])
    kp_uv = rng.uniform(0, 700, (n + extra, 2)).astype(np.float32); kp_oct = rng.integers(0, 8, n + extra).astype(np.int32)
    mp_of_kp = np.full(n + extra, -1, np.int32)
    kp_uv[slot] = d["uv"]; kp_oct[slot] = octave; mp_of_kp[slot] = np.arange(n)
    a, b = both.run(H.run_pose_optimization, sc, kp_uv, kp_oct, mp_of_kp, T32, d["intr"])
    same_out(a, b)


def sim3_scene(d, n, seed, drop=0.1, nbad_div=10, unseen_frac=0.05):
    """two keyframes at the identity pose holding the pairs of a synthetic Sim3 alignment problem (so that R*P + t in f32 is P)"""
    rng = np.random.default_rng(seed + 1)
    o1 = rng.integers(0, 8, n).astype(np.int32); o2 = rng.integers(0, 8, n).astype(np.int32)
    P1, P2 = np.float32(d["P1c"]), np.float32(d["P2c"])
    I = np.eye(4, dtype=np.float32)
    match1 = (n + np.arange(n)).astype(np.int32); match1[rng.random(n) < drop] = -1
    bad = np.zeros(2 * n, np.uint8); bad[rng.permutation(2 * n)[:n // nbad_div]] = 1
    unseen = rng.random(n) < unseen_frac
    obs_cnt = np.r_[np.ones(n, np.int32), (~unseen).astype(np.int32)]
    sc = dict(kf_uid=np.arange(2).astype(np.int64), kf_id=np.stack([np.arange(2), np.zeros(2, np.int64)], 1), kf_bad=np.zeros(2, np.uint8),
              kf_Tcw=np.stack([I, I]), kf_intr=np.stack([np.float32(d["K1"]), np.float32(d["K2"])]), kp_ptr=np.array([0, n, 2 * n], np.int32),
              kp_uv=np.r_[np.float32(d["uv1"]), np.float32(d["uv2"])], kp_octave=np.r_[o1, o2], inv_level_sigma2=H.sm.INV_LEVEL_SIGMA2, kf_parent=None,
              loop_ptr=None, loop_kf=None, cov_ptr=None, cov_kf=None, cov_w=None, mp_uid=np.arange(2 * n).astype(np.int64),
              mp_id=np.stack([np.arange(2 * n), np.zeros(2 * n, np.int64)], 1), mp_bad=bad, mp_pos=np.r_[P1, P2], mp_ref=np.zeros(2 * n, np.int32),
              obs_ptr=np.concatenate([[0], np.cumsum(obs_cnt)]).astype(np.int32), obs_kf=np.r_[np.zeros(n, np.int32), np.ones((~unseen).sum(), np.int32)],
              obs_idx=np.r_[np.arange(n), np.arange(n)[~unseen]].astype(np.int32), origin=0, map_id=0)
    return sc, match1


@pytest.mark.parametrize("n,seed,fix,frac", [(120, 12, False, 0.2), (120, 12, True, 0.2), (12, 5, False, 0.6), (60, 7, False, 0.0)])
def test_reference_optimize_sim3(ho, both, n, seed, fix, frac):
    d = synth.make_sim3_opt(n=n, seed=seed, fix_scale=fix, outlier_frac=frac)
    sc, match1 = sim3_scene(d, n, seed)
    a, b = both.run(H.run_optimize_sim3, sc, 0, 1, match1, d["S12_0"], d["th2"], fix)
    same_out(a, b)


@pytest.mark.parametrize("seed", range(8))
def test_reference_optimize_sim3_sweep(ho, both, seed):
    rng = np.random.default_rng(500 + seed)
    n = int(rng.choice([11, 25, 70, 200])); fix = bool(rng.random() < 0.5)
    d = synth.make_sim3_opt(n=n, seed=60 + seed, fix_scale=fix, outlier_frac=float(rng.choice([0.0, 0.15, 0.4, 0.7])))
    sc, match1 = sim3_scene(d, n, seed, drop=float(rng.choice([0.0, 0.2])), unseen_frac=float(rng.choice([0.0, 0.1])))
    a, b = both.run(H.run_optimize_sim3, sc, 0, 1, match1, d["S12_0"], d["th2"], fix)
    same_out(a, b)


@pytest.mark.parametrize("fix_scale", [False, True])
def test_reference_essential_graph(ho, both, fix_scale):
    K = 40
    sc, cov, loops = essential_scene(ho, K=K, seed=1)
    loop_kf, cur_kf = 2, K - 1
    conn = {cur_kf: [loop_kf, 4], loop_kf: [cur_kf], 4: [cur_kf], K - 2: [3]}
    same_out(*both.run(H.run_essential_graph, sc, loop_kf, cur_kf, conn, fix_scale))
    # loop closure variant with corrected / non-corrected maps and tagged points
    O = ho.Pieces("oracle"); R = ho.Pieces("ref")
    rng = np.random.default_rng(5)
    near = [K - 1, K - 2, K - 3]
    fix = O.vec("sim3_exp", 8, np.r_[rng.normal(0, 0.02, 3), rng.normal(0, 0.1, 3), 0.0 if fix_scale else 0.05])
    non, cor = [], []
    for k in near:
        T = sc["kf_Tcw"][k].astype(np.float64)
        non.append(R.vec("sim3_from_Rt", 8, T[:3, :3].ravel(), T[:3, 3], [1.0])); cor.append(O.vec("sim3_mul", 8, non[-1], fix))
    mp_corr = np.where(np.random.default_rng(6).random(300) < 0.2, sc["kf_uid"][K - 2], -1).astype(np.int32)
    a, b = both.run(H.run_essential_graph, sc, loop_kf, cur_kf, {cur_kf: [loop_kf, 4], loop_kf: [cur_kf], 4: [cur_kf]}, fix_scale, loop_closure=True,
                    corr=(near, np.stack(cor), np.stack(non)), mp_corr_ref=mp_corr)
    same_out(a, b)
    assert np.abs(a["kf_Tcw"] - sc["kf_Tcw"]).max() > 1e-3


@pytest.mark.parametrize("seed", range(6))
def test_reference_sweep(ho, both, seed):
    """more scenes, same claim: random bad keyframes / points, iteration counts, kernels on and off, different agents' keyframes in one map"""
    rng = np.random.default_rng(100 + seed)
    p = synth.make_config("small")
    client = (np.arange(p.K) >= int(p.K * rng.uniform(0.3, 0.8))).astype(np.int64)   # two agents' keyframes: mId = (k, client), mUniqueId by GetID
    sc = H.scene_from_problem(p, ho, seed=seed, map_id=0, bad_kf=float(rng.choice([0.0, 0.05, 0.2])), bad_mp=float(rng.choice([0.0, 0.1, 0.3])),
                              client_of_kf=client)
    sc["kf_bad"][0] = 0; sc["kf_id"][0] = (0, 0); sc["kf_uid"][0] = 0
    its = int(rng.choice([1, 3, 6, 12])); robust = bool(rng.random() < 0.7)
    loop = (int(rng.integers(1, 30)), int(rng.integers(0, 2))) if rng.random() < 0.5 else (0, 0)
    same_out(*both.run(H.run_gba, sc, 0, its, robust, loop))
    d = synth.make_pose_opt(n=int(rng.choice([15, 80, 400])), seed=200 + seed, outlier_frac=float(rng.choice([0.0, 0.2, 0.45])),
                            pose_noise=(0.05, 0.2) if seed % 2 else (0.03, 0.08))
    n = len(d["uv"]); octave = rng.integers(0, 8, n).astype(np.int32)
    scp = dict(kf_uid=np.zeros(1, np.int64), kf_id=np.zeros((1, 2), np.int64), kf_bad=np.zeros(1, np.uint8), kf_Tcw=np.eye(4, dtype=np.float32)[None],
               kf_intr=np.float32(d["intr"])[None], kp_ptr=np.zeros(2, np.int32), kp_uv=np.zeros((0, 2), np.float32), kp_octave=np.zeros(0, np.int32),
               inv_level_sigma2=H.sm.INV_LEVEL_SIGMA2, kf_parent=None, loop_ptr=None, loop_kf=None, cov_ptr=None, cov_kf=None, cov_w=None,
               mp_uid=np.arange(n).astype(np.int64), mp_id=np.stack([np.arange(n), np.zeros(n, np.int64)], 1), mp_bad=np.zeros(n, np.uint8),
               mp_pos=np.float32(d["Xw"]), mp_ref=np.zeros(n, np.int32), obs_ptr=np.zeros(n + 1, np.int32), obs_kf=np.zeros(0, np.int32),
               obs_idx=np.zeros(0, np.int32), origin=0, map_id=0)
    same_out(*both.run(H.run_pose_optimization, scp, np.float32(d["uv"]), octave, np.arange(n).astype(np.int32), ho.pose_to_Tcw_f32(d["Tcw0"]), d["intr"]))


def test_reference_vertex_order_is_the_only_difference(ho, both):
    """g2o numbers its vertices by id (mUniqueId); the shim's rows follow the map's iteration order.  When the two orders differ — here the
    keyframes of two agents interleaved — the same system is eliminated in another order and the f64 results part in the last bits:
    the f32 poses written back agree to an ulp instead of exactly.  (On the device the sums are reordered anyway.)"""
    rng = np.random.default_rng(105)
    p = synth.make_config("small")
    client = (rng.random(p.K) < 0.4).astype(np.int64)
    sc = H.scene_from_problem(p, ho, seed=5, map_id=0, bad_kf=0.05, bad_mp=0.3, client_of_kf=client)
    sc["kf_bad"][0] = 0; sc["kf_id"][0] = (0, 0); sc["kf_uid"][0] = 0
    a, b = both.run(H.run_gba, sc, 0, 12, True, (0, 0))
    for k in a:
        if a[k].dtype == np.float32:
            assert np.abs(a[k] - b[k]).max() <= 2 * np.spacing(np.float32(np.abs(b[k]).max())), k
        else:
            assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("name", ["ba_tiny.npz", "ba_small.npz"])
def test_committed_ba_fixtures_are_what_the_reference_computes(ho, both, name):
    """tests/golden/ba_*.npz were written by the oracle (tests/golden/make_golden.py); the GPU suite compares the device path with them
    on the GPU box, where /root/reference does not exist.  Here the same problems go through the reference's OWN Optimizer::MapFusionGBA
    (oracle/_ref/liboptimizer_ref.so): the poses and points it writes back are the fixture's, rounded to f32 as the reference stores them."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))
    p = synth.BAProblem(poses=g["in_poses"], intr=g["in_intr"], fixed=g["in_fixed"], points=g["in_points"], obs_kf=g["in_obs_kf"], obs_mp=g["in_obs_mp"],
                        obs_uv=g["in_obs_uv"], obs_w=g["in_obs_w"])
    assert abs(float(g["huber_delta"]) - api.HUBER_GBA) < 1e-15 and p.fixed.sum() == 1
    sc = H.scene_from_problem(p, ho, seed=0, map_id=0, keep_weights=True)
    sc["origin"] = int(np.flatnonzero(p.fixed)[0])
    # the reference starts from the f32 poses a KeyFrame holds: only exact when the fixture's poses survive the f32 round trip
    start = np.stack([ho.pose_from_Tcw_f32(T) for T in sc["kf_Tcw"]])
    H.use_reference(True)
    try:
        out = H.run_gba(sc, 0, int(g["iterations"]), True, (0, 0))
    finally:
        H.use_reference(False)
    p2 = p.copy(); p2.poses = start; p2.points = p.points.astype(np.float32).astype(np.float64)
    want = ho.ba_solve(p2, iterations=int(g["iterations"]), robust=True, huber_delta=api.HUBER_GBA)
    assert np.array_equal(out["kf_Tcw"], np.stack([ho.pose_to_Tcw_f32(q) for q in want["poses"]]))
    assert np.array_equal(out["mp_pos"], want["points"].astype(np.float32))
    # and the fixture itself (f64 start) lands within f32 resolution of that
    Tfix = np.stack([ho.pose_to_Tcw_f32(q) for q in g["poses"]])
    assert np.abs(out["kf_Tcw"] - Tfix).max() <= 1e-5 * max(1.0, np.abs(Tfix).max())
    assert np.abs(out["mp_pos"] - g["points"]).max() <= 1e-5 * np.abs(g["points"]).max()


def test_reference_error_conventions(ho, both):
    """the fatal cases of SURVEY.md 8(b): both implementations throw estd::infrastructure_ex and leave the map alone"""
    p = synth.make_config("tiny")
    sc = H.scene_from_problem(p, ho, seed=1, map_id=0)
    no_origin = dict(sc, origin=-1)                                 # pMap->mvpKeyFrameOrigins.empty()  (S/Optimizer.cpp:663-667)
    for ref in (False, True):
        H.use_reference(ref)
        try:
            out = H.run_gba(no_origin, 0, 5, True, (0, 0), want_rc=-1)
        finally:
            H.use_reference(False)
        assert not out["kf_set_pose"].any()
    big = dict(sc); big["kf_id"] = sc["kf_id"].copy(); big["kf_id"][1, 0] = 1000000     # keyframe id >= IDRANGE  (S/Optimizer.cpp:68-72)
    for ref in (False, True):
        H.use_reference(ref)
        try:
            H.run_gba(big, 1, 5, True, (0, 0), want_rc=-1)
        finally:
            H.use_reference(False)


def test_shim_is_reentrant(ho):
    """SURVEY.md 8(b) threading: the entry points are called concurrently (per-agent LocalBA threads, per-map GBA threads).  Four threads
    run MapFusionGBA / LocalBundleAdjustmentClient through the shim at once on their own scenes; each must get what it gets alone."""
    import threading
    jobs = []
    for seed in range(4):
        p = synth.make_config("small")
        sc = H.scene_from_problem(p, ho, seed=40 + seed, map_id=0, bad_kf=0.05 * seed, bad_mp=0.1)
        sc["kf_bad"][0] = 0
        jobs.append(sc)
    H.use_reference(False)
    alone = [H.run_gba(sc, 0, 6, True, (0, 0)) for sc in jobs]
    got = [None] * len(jobs); err = []

    def work(i):
        try:
            for _ in range(3):
                got[i] = H.run_gba(jobs[i], 0, 6, True, (0, 0))
        except Exception as e:          # noqa: BLE001
            err.append(e)
    th = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for t in th: t.start()
    for t in th: t.join()
    assert not err
    for a, b in zip(alone, got):
        same_out(a, b)


@pytest.mark.parametrize("seed", range(8))
def test_reference_local_ba_sweep(ho, both, seed):
    """LocalBundleAdjustmentClient has the most selection logic (local window, local points, fixed observers, two rounds, erasures): random
    centres, covisibility lists, bad keyframes / points and map sizes through both implementations"""
    rng = np.random.default_rng(300 + seed)
    p = synth.make_config("cfg2", P=int(rng.choice([150, 400, 900])))
    sc = H.scene_from_problem(p, ho, seed=seed, map_id=0, bad_kf=float(rng.choice([0.0, 0.1])), bad_mp=float(rng.choice([0.0, 0.1, 0.3])))
    center = int(rng.integers(0, p.K))
    sc["kf_bad"][center] = 0
    ncov = int(rng.integers(0, 14))
    covis = [int(k) for k in rng.permutation(p.K) if k != center][:ncov]
    cov_ptr = np.zeros(p.K + 1, np.int32); cov_ptr[center + 1:] = len(covis)
    sc.update(cov_ptr=cov_ptr, cov_kf=np.array(covis, np.int32), cov_w=np.full(len(covis), 200, np.int32))
    a, b = both.run(H.run_local_ba, sc, center, bool(seed % 2))
    same_out(a, b)
    assert a["kf_set_pose"][center] == 1


@pytest.mark.parametrize("seed", range(6))
def test_reference_essential_graph_sweep(ho, both, seed):
    """random spanning trees, earlier loop edges, covisibility lists with weights either side of the threshold, random new loop connections,
    free and fixed scale, both variants — shim vs the reference's OptimizeEssentialGraph*"""
    rng = np.random.default_rng(400 + seed)
    K = int(rng.choice([12, 30, 55]))
    sc, cov, loops = essential_scene(ho, K=K, seed=50 + seed)
    sc["kf_parent"] = np.array([-1] + [int(rng.integers(0, k)) for k in range(1, K)], np.int32)        # a random tree rooted at keyframe 0
    pairs = [(int(rng.integers(2, K)), int(rng.integers(0, 2))) for _ in range(int(rng.integers(0, 4)))]
    le = {k: set() for k in range(K)}
    for a, b in pairs:
        if a != b:
            le[a].add(b); le[b].add(a)
    sc["loop_ptr"] = np.concatenate([[0], np.cumsum([len(le[k]) for k in range(K)])]).astype(np.int32)
    sc["loop_kf"] = np.array([j for k in range(K) for j in sorted(le[k])], np.int32)
    loop_kf, cur_kf = int(rng.integers(0, K // 2)), K - 1
    conn = {cur_kf: sorted({loop_kf, int(rng.integers(0, K - 1))}), loop_kf: [cur_kf]}
    fix = bool(seed % 2)
    same_out(*both.run(H.run_essential_graph, sc, loop_kf, cur_kf, conn, fix))
    O = ho.Pieces("oracle"); R = ho.Pieces("ref")
    near = sorted(set([cur_kf] + [int(k) for k in rng.integers(0, K, 3)]))
    d = O.vec("sim3_exp", 8, np.r_[rng.normal(0, 0.03, 3), rng.normal(0, 0.1, 3), 0.0 if fix else rng.normal(0, 0.05)])
    non, cor = [], []
    for k in near:
        T = sc["kf_Tcw"][k].astype(np.float64)
        non.append(R.vec("sim3_from_Rt", 8, T[:3, :3].ravel(), T[:3, 3], [1.0])); cor.append(O.vec("sim3_mul", 8, non[-1], d))
    tag = np.where(rng.random(300) < 0.15, sc["kf_uid"][near[0]], -1).astype(np.int32)
    same_out(*both.run(H.run_essential_graph, sc, loop_kf, cur_kf, conn, fix, loop_closure=True, corr=(near, np.stack(cor), np.stack(non)), mp_corr_ref=tag))
