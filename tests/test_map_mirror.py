"""CPU suite, SURVEY.md §8(f) rank 1: the persistent flat mirror of the map (ccm_mirror_*, host code of the product library).
A random history of the changes a running server makes — keyframes and points inserted, moved, flagged bad, erased, observations
added / re-measured / erased — is applied to the mirror and to a plain Python model of the map; after every burst the problem the
mirror hands out must equal, array for array, the one a from-scratch flattening of the model gives under MapFusionGBA's selection
rules (S/Optimizer.cpp:693-786).  Value-only bursts must not trigger a rebuild, and must still be visible in the arrays."""
import numpy as np
import pytest

from ccm_slam_b200 import api


def rand_T(rng):
    w = rng.normal(0, 0.5, 3); th = np.linalg.norm(w); k = w / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    T = np.eye(4); T[:3, :3] = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx; T[:3, 3] = rng.normal(0, 5, 3)
    return T.astype(np.float32)


class Model:
    """insertion-ordered dicts; erase + insert again moves to the end, an update keeps the place"""

    def __init__(self):
        self.kf, self.mp, self.obs = {}, {}, {}

    def flatten(self, max_uid, fixed, min_edges=2):
        """MapFusionGBA (min_edges = 2, S/Optimizer.cpp:722-740: observations.size() < 2 || nEdges < 2 skips the point) or
        BundleAdjustmentClient (min_edges = 1); observations grouped by point row, insertion order inside a group"""
        rows = {}; poses = []; intr = []; fx = []
        for uid, k in self.kf.items():
            if k["bad"] or uid > max_uid:
                continue
            rows[uid] = len(poses); poses.append(k["T"]); intr.append(k["intr"]); fx.append(uid in fixed)
        nobs, nedges = {}, {}
        for (m, k), o in self.obs.items():
            if m in self.mp and not self.mp[m]["bad"]:
                nobs[m] = nobs.get(m, 0) + 1
                if k in rows:
                    nedges[m] = nedges.get(m, 0) + 1
        prow = {}; pts = []
        for uid, p in self.mp.items():
            if p["bad"] or nobs.get(uid, 0) < min_edges or nedges.get(uid, 0) < min_edges:
                continue
            prow[uid] = len(pts); pts.append(p["x"])
        edges = [(m, k, o) for (m, k), o in self.obs.items() if k in rows and m in prow]
        edges.sort(key=lambda e: prow[e[0]])                       # stable: insertion order inside a point's group
        T = np.array(poses, np.float32).reshape(-1, 16)
        return dict(poses=api.poses_from_Tcw_f32(T) if len(T) else np.zeros((0, 7)), intr=np.array(intr, np.float32).astype(np.float64).reshape(-1, 4),
                    fixed=np.array(fx, np.uint8), points=np.array(pts, np.float32).astype(np.float64).reshape(-1, 3),
                    obs_kf=np.array([rows[k] for _, k, _ in edges], np.int32), obs_mp=np.array([prow[m] for m, _, _ in edges], np.int32),
                    obs_uv=np.array([o[:2] for _, _, o in edges], np.float32).reshape(-1, 2), obs_w=np.array([o[2] for _, _, o in edges], np.float32)), \
            np.array(list(rows), np.uint64), np.array(list(prow), np.uint64)


def same(got, want):
    a, ku, mu = got; b, rku, rmu = want
    assert np.array_equal(ku, rku) and np.array_equal(mu, rmu)
    for key in b:
        assert np.array_equal(a[key], b[key]), key


def burst(rng, M, mir, n, next_uid, structural=True):
    for _ in range(n):
        op = rng.choice(["kf", "mp", "obs", "move_kf", "move_mp", "bad_kf", "bad_mp", "del_kf", "del_mp", "del_obs", "remeasure"] if structural else ["move_kf", "move_mp"],
                        p=[0.08, 0.2, 0.3, 0.1, 0.1, 0.02, 0.03, 0.02, 0.05, 0.07, 0.03] if structural else [0.5, 0.5])
        if op == "kf" or (op in ("move_kf", "bad_kf", "del_kf") and not M.kf):
            uid = next_uid[0]; next_uid[0] += int(rng.integers(1, 3))
            k = dict(T=rand_T(rng), intr=rng.uniform(300, 600, 4).astype(np.float32), bad=False)
            M.kf[uid] = k; mir.set_keyframe(uid, k["T"], k["intr"])
        elif op == "mp" or (op in ("move_mp", "bad_mp", "del_mp") and not M.mp):
            uid = next_uid[1]; next_uid[1] += 1
            p = dict(x=rng.normal(0, 10, 3).astype(np.float32), bad=False)
            M.mp[uid] = p; mir.set_point(uid, p["x"])
        elif op == "obs" and M.kf and M.mp:
            k = int(rng.choice(list(M.kf))); m = int(rng.choice(list(M.mp)))
            o = (np.float32(rng.uniform(0, 700)), np.float32(rng.uniform(0, 400)), np.float32(1.2 ** -int(rng.integers(0, 8))))
            M.obs[(m, k)] = o; mir.set_observation(k, m, *o)        # an existing pair keeps its place (dict update)
        elif op == "remeasure" and M.obs:
            (m, k) = list(M.obs)[int(rng.integers(0, len(M.obs)))]
            o = (np.float32(rng.uniform(0, 700)), np.float32(rng.uniform(0, 400)), np.float32(1.0))
            M.obs[(m, k)] = o; mir.set_observation(k, m, *o)
        elif op == "move_kf":
            uid = int(rng.choice(list(M.kf))); M.kf[uid]["T"] = rand_T(rng); mir.set_keyframe(uid, M.kf[uid]["T"], None, M.kf[uid]["bad"])
        elif op == "move_mp":
            uid = int(rng.choice(list(M.mp))); M.mp[uid]["x"] = rng.normal(0, 10, 3).astype(np.float32); mir.set_point(uid, M.mp[uid]["x"], M.mp[uid]["bad"])
        elif op == "bad_kf":
            uid = int(rng.choice(list(M.kf))); M.kf[uid]["bad"] = not M.kf[uid]["bad"]; mir.set_keyframe(uid, M.kf[uid]["T"], None, M.kf[uid]["bad"])
        elif op == "bad_mp":
            uid = int(rng.choice(list(M.mp))); M.mp[uid]["bad"] = not M.mp[uid]["bad"]; mir.set_point(uid, M.mp[uid]["x"], M.mp[uid]["bad"])
        elif op == "del_kf":
            uid = int(rng.choice(list(M.kf))); del M.kf[uid]; mir.erase_keyframe(uid)
            for key in [key for key in M.obs if key[1] == uid]:
                del M.obs[key]
        elif op == "del_mp":
            uid = int(rng.choice(list(M.mp))); del M.mp[uid]; mir.erase_point(uid)
            for key in [key for key in M.obs if key[0] == uid]:
                del M.obs[key]
        elif op == "del_obs" and M.obs:
            (m, k) = list(M.obs)[int(rng.integers(0, len(M.obs)))]
            del M.obs[(m, k)]; mir.erase_observation(k, m)


@pytest.mark.parametrize("seed", range(6))
def test_random_history(seed):
    rng = np.random.default_rng(seed)
    M = Model(); mir = api.MapMirror(); nxt = [0, 0]; most = [0, 0]
    min_edges = 1 + seed % 2
    mir.set_min_edges(min_edges)
    for rnd in range(12):
        burst(rng, M, mir, int(rng.integers(50, 400)), nxt)
        max_uid = int(rng.integers(0, max(nxt[0], 1) + 2)) if rnd % 3 == 2 else 10 ** 9          # sometimes a cut like MapFusionGBA's maxKFid
        fixed = {0} if rnd % 2 == 0 else {int(u) for u in list(M.kf)[:2]}
        _, a, ku, mu = mir.problem(max_uid, sorted(fixed))
        same((a, ku, mu), M.flatten(max_uid, fixed, min_edges))
        # value-only burst: no rebuild, still exact
        r0 = mir.rebuilds()
        _, a, ku, mu = mir.problem(max_uid, sorted(fixed)); assert mir.rebuilds() == r0               # the same question again: cached
        burst(rng, M, mir, 60, nxt, structural=False)
        _, a, ku, mu = mir.problem(max_uid, sorted(fixed))
        assert mir.rebuilds() == r0
        same((a, ku, mu), M.flatten(max_uid, fixed, min_edges))
        most = [max(most[0], len(a["obs_kf"])), max(most[1], len(ku))]
    assert most[0] > 50 and most[1] > 5
    mir.close()


def test_empty_and_errors():
    mir = api.MapMirror()
    prob, a, ku, mu = mir.problem(10, [])
    assert prob.K == prob.P == prob.E == 0
    with pytest.raises(api.CCMError):
        mir.set_observation(1, 2, 0.0, 0.0, 1.0)                 # neither end exists
    with pytest.raises(api.CCMError):
        mir.set_keyframe(5, np.eye(4, dtype=np.float32))          # a new keyframe needs intrinsics
    mir.erase_keyframe(99); mir.erase_point(99); mir.erase_observation(1, 2)   # erasing what is not there is a no-op, as in the reference's containers
    mir.close()


def test_mirror_problem_solves_like_the_flattened_one(oracle):
    """a synthetic BA problem loaded through the mirror row by row gives ccm_ba_solve's oracle the same arrays as the problem itself"""
    from ccm_slam_b200 import synth
    p = synth.make_config("small")
    T = api.poses_to_Tcw_f32(p.poses)
    mir = api.MapMirror()
    for k in range(p.K):
        mir.set_keyframe(k, T[k], p.intr[k].astype(np.float32))
    pts = p.points.astype(np.float32)
    for i in range(p.P):
        mir.set_point(i, pts[i])
    for e in range(p.E):
        mir.set_observation(int(p.obs_kf[e]), int(p.obs_mp[e]), p.obs_uv[e, 0], p.obs_uv[e, 1], p.obs_w[e])
    _, a, ku, mu = mir.problem(10 ** 9, np.flatnonzero(p.fixed))
    assert np.array_equal(a["obs_kf"], p.obs_kf) and np.array_equal(a["obs_mp"], p.obs_mp) and np.array_equal(a["obs_uv"], p.obs_uv)
    assert np.array_equal(a["fixed"], p.fixed) and np.array_equal(a["poses"], api.poses_from_Tcw_f32(T)) and np.array_equal(a["points"], pts.astype(np.float64))
    mir.close()


def test_compaction_after_mass_erasure():
    """most of the map erased (keyframe culling, a map reset): dead slots are dropped, order and content survive, and what was erased
    can come back under the same ids"""
    rng = np.random.default_rng(77)
    M = Model(); mir = api.MapMirror(); nxt = [0, 0]
    burst(rng, M, mir, 1500, nxt)
    for uid in list(M.kf)[::2] + list(M.kf)[1::4]:
        if uid in M.kf:
            del M.kf[uid]; mir.erase_keyframe(uid)
            for key in [key for key in M.obs if key[1] == uid]:
                del M.obs[key]
    for uid in list(M.mp)[: len(M.mp) * 3 // 4]:
        del M.mp[uid]; mir.erase_point(uid)
        for key in [key for key in M.obs if key[0] == uid]:
            del M.obs[key]
    _, a, ku, mu = mir.problem(10 ** 9, [0])
    same((a, ku, mu), M.flatten(10 ** 9, {0}))
    # an erased id comes back: new slot, end of the order; old observations of that id stay gone
    back = 0 if 0 not in M.kf else next(u for u in range(10 ** 6) if u not in M.kf)
    k = dict(T=rand_T(rng), intr=np.float32([400, 400, 300, 200]), bad=False)
    M.kf[back] = k; mir.set_keyframe(back, k["T"], k["intr"])
    burst(rng, M, mir, 400, nxt)
    _, a, ku, mu = mir.problem(10 ** 9, [back])
    same((a, ku, mu), M.flatten(10 ** 9, {back}))
    assert len(ku) > 10 and len(a["obs_kf"]) > 20
    mir.close()


def test_row_order_only_reorders_the_sums(oracle):
    """the mirror's rows follow insertion order, the reference's follow ids / pointer values: the same BA, summed in another order.
    A problem fed to the mirror in shuffled order solves (CPU oracle) to the same LM trajectory and the same estimates, matched by uid."""
    from ccm_slam_b200 import synth
    p = synth.make_config("small")
    T = api.poses_to_Tcw_f32(p.poses); pts = p.points.astype(np.float32)
    base = synth.BAProblem(poses=api.poses_from_Tcw_f32(T), intr=p.intr, fixed=p.fixed, points=pts.astype(np.float64), obs_kf=p.obs_kf, obs_mp=p.obs_mp,
                           obs_uv=p.obs_uv, obs_w=p.obs_w)
    rng = np.random.default_rng(5)
    mir = api.MapMirror()
    for k in rng.permutation(p.K):
        mir.set_keyframe(int(k), T[k], p.intr[k].astype(np.float32))
    for i in rng.permutation(p.P):
        mir.set_point(int(i), pts[i])
    for e in rng.permutation(p.E):
        mir.set_observation(int(p.obs_kf[e]), int(p.obs_mp[e]), p.obs_uv[e, 0], p.obs_uv[e, 1], p.obs_w[e])
    _, a, ku, mu = mir.problem(10 ** 9, np.flatnonzero(p.fixed))
    mir.close()
    assert sorted(ku) == list(range(p.K)) and sorted(mu) == list(range(p.P)) and not np.array_equal(ku, np.arange(p.K))
    shuffled = synth.BAProblem(poses=a["poses"], intr=a["intr"], fixed=a["fixed"], points=a["points"], obs_kf=a["obs_kf"], obs_mp=a["obs_mp"],
                               obs_uv=a["obs_uv"], obs_w=a["obs_w"])
    r0 = oracle.ba_solve(base, iterations=8); r1 = oracle.ba_solve(shuffled, iterations=8)
    assert len(r0["trace"]) == len(r1["trace"]) and np.allclose(r0["trace"][:, 2], r1["trace"][:, 2], rtol=1e-9)
    assert np.abs(r1["poses"][np.argsort(ku)] - r0["poses"]).max() < 1e-8
    assert np.abs(r1["points"][np.argsort(mu)] - r0["points"]).max() < 1e-7
