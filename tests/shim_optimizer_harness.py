"""ctypes harness for oracle/_ref/liboptimizer_shim.so (test infrastructure): shim/Optimizer_shim.cpp behind the reference's own
cslam::Optimizer interface, on stand-in Map / KeyFrame / MapPoint / Frame objects built from flat arrays (oracle/ref_optimizer_wrap.cpp),
with the device entry points doubled by the CPU oracle (oracle/ccm_device_double.cpp)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ccm_slam_b200 import synth
from ccm_slam_b200 import synth_match as sm

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "..", "oracle", "_ref", "liboptimizer_shim.so")
_SO_GPU = os.path.join(_HERE, "..", "oracle", "_ref", "liboptimizer_shim_gpu.so")   # the same shim over the real device entry points
_LIB = None
_GPU = False
VP = C.c_void_p


class Scene(C.Structure):
    _fields_ = [("K", C.c_int32), ("kf_uid", VP), ("kf_id", VP), ("kf_bad", VP), ("kf_Tcw", VP), ("kf_intr", VP), ("kp_ptr", VP), ("kp_uv", VP),
                ("kp_octave", VP), ("inv_level_sigma2", VP), ("nlevels", C.c_int32), ("kf_parent", VP), ("loop_ptr", VP), ("loop_kf", VP),
                ("cov_ptr", VP), ("cov_kf", VP), ("cov_w", VP), ("P", C.c_int32), ("mp_uid", VP), ("mp_id", VP), ("mp_bad", VP), ("mp_pos", VP),
                ("mp_ref", VP), ("obs_ptr", VP), ("obs_kf", VP), ("obs_idx", VP), ("origin", C.c_int32), ("map_id", C.c_int64)]


class Out(C.Structure):
    _fields_ = [(n, VP) for n in ("kf_Tcw", "kf_TcwGBA", "kf_gba_tag", "kf_set_pose", "mp_pos", "mp_posGBA", "mp_gba_tag", "mp_set_pos",
                                  "mp_update_normal", "mp_n_obs", "kf_n_erased")]


def lib():
    global _LIB
    if _LIB is None:
        from oracle import pyoracle
        so = _SO_REF if _REF else (_SO_GPU if _GPU else _SO)
        if pyoracle.build_ref() is None or not os.path.exists(so):
            return None
        _LIB = C.CDLL(so)
    return _LIB


_SO_REF = os.path.join(_HERE, "..", "oracle", "_ref", "liboptimizer_ref.so")   # the reference's OWN Optimizer.cpp behind the same wrapper
_REF = False


def use_reference(on):
    """switch the harness to the library built from the reference's own cslam/src/Optimizer.cpp + g2o core (oracle-backed linear solvers)"""
    global _LIB, _REF
    _LIB, _REF = None, bool(on)


def use_device(on):
    """switch the harness between the CPU-doubled library and the one linked against the product's device entry points"""
    global _LIB, _GPU
    _LIB, _GPU = None, bool(on)


def _p(a):
    return None if a is None else a.ctypes.data_as(VP)


SCENE_ARRAYS = dict(kf_uid=np.int64, kf_id=np.int64, kf_bad=np.uint8, kf_Tcw=np.float32, kf_intr=np.float32, kp_ptr=np.int32, kp_uv=np.float32,
                    kp_octave=np.int32, inv_level_sigma2=np.float32, kf_parent=np.int32, loop_ptr=np.int32, loop_kf=np.int32, cov_ptr=np.int32,
                    cov_kf=np.int32, cov_w=np.int32, mp_uid=np.int64, mp_id=np.int64, mp_bad=np.uint8, mp_pos=np.float32, mp_ref=np.int32,
                    obs_ptr=np.int32, obs_kf=np.int32, obs_idx=np.int32)


def c_scene(sc, keep):
    a = {k: (None if sc.get(k) is None else np.ascontiguousarray(sc[k], dt)) for k, dt in SCENE_ARRAYS.items()}
    keep.append(a)
    K, P = len(a["kf_uid"]), len(a["mp_uid"])
    return Scene(K, _p(a["kf_uid"]), _p(a["kf_id"]), _p(a["kf_bad"]), _p(a["kf_Tcw"]), _p(a["kf_intr"]), _p(a["kp_ptr"]), _p(a["kp_uv"]),
                 _p(a["kp_octave"]), _p(a["inv_level_sigma2"]), len(a["inv_level_sigma2"]), _p(a["kf_parent"]), _p(a["loop_ptr"]), _p(a["loop_kf"]),
                 _p(a["cov_ptr"]), _p(a["cov_kf"]), _p(a["cov_w"]), P, _p(a["mp_uid"]), _p(a["mp_id"]), _p(a["mp_bad"]), _p(a["mp_pos"]),
                 _p(a["mp_ref"]), _p(a["obs_ptr"]), _p(a["obs_kf"]), _p(a["obs_idx"]), int(sc["origin"]), int(sc["map_id"]))


def new_out(K, P):
    o = dict(kf_Tcw=np.zeros((K, 4, 4), np.float32), kf_TcwGBA=np.zeros((K, 4, 4), np.float32), kf_gba_tag=np.zeros((K, 2), np.int64),
             kf_set_pose=np.zeros(K, np.int32), mp_pos=np.zeros((P, 3), np.float32), mp_posGBA=np.zeros((P, 3), np.float32),
             mp_gba_tag=np.zeros((P, 2), np.int64), mp_set_pos=np.zeros(P, np.int32), mp_update_normal=np.zeros(P, np.int32),
             mp_n_obs=np.zeros(P, np.int32), kf_n_erased=np.zeros(K, np.int32))
    return o, Out(*[_p(o[n]) for n, _ in Out._fields_])


def scene_from_problem(p, oracle, seed=0, map_id=0, bad_kf=0.0, bad_mp=0.0, client_of_kf=None, keep_weights=False):
    """a stand-in map holding the flat BA problem p: keyframe k <- pose k (as the f32 Tcw the reference stores), one keypoint per observation"""
    rng = np.random.default_rng(seed)
    K, P = p.K, p.P
    Tcw = np.stack([oracle.pose_to_Tcw_f32(q) for q in p.poses])
    order = np.lexsort((p.obs_kf, p.obs_mp))                       # observations of a point, by keyframe index
    okf, omp, ouv = p.obs_kf[order], p.obs_mp[order], p.obs_uv[order]
    octave = rng.integers(0, 8, len(order)).astype(np.int32)
    table = sm.INV_LEVEL_SIGMA2
    if keep_weights:                                               # the problem's own weights: they become the level table, the octaves index it
        table = np.sort(np.unique(p.obs_w))[::-1].astype(np.float32)
        octave = np.array([int(np.flatnonzero(table == w)[0]) for w in p.obs_w[order]], np.int32)
    kp_count = np.zeros(K, np.int64); obs_idx = np.zeros(len(order), np.int32)
    for e, k in enumerate(okf):
        obs_idx[e] = kp_count[k]; kp_count[k] += 1
    kp_ptr = np.concatenate([[0], np.cumsum(kp_count)]).astype(np.int32)
    kp_uv = np.zeros((kp_ptr[-1], 2), np.float32); kp_oct = np.zeros(kp_ptr[-1], np.int32)
    kp_uv[kp_ptr[okf] + obs_idx] = ouv; kp_oct[kp_ptr[okf] + obs_idx] = octave
    obs_ptr = np.concatenate([[0], np.cumsum(np.bincount(omp, minlength=P))]).astype(np.int32)
    client = np.zeros(K, np.int64) if client_of_kf is None else np.asarray(client_of_kf, np.int64)
    kf_id = np.stack([np.arange(K), client], 1).astype(np.int64)
    kf_uid = (1000000 * client + np.arange(K)).astype(np.int64)    # Optimizer::GetID
    mp_id = np.stack([np.arange(P), np.zeros(P, np.int64)], 1)
    ref = np.array([okf[obs_ptr[j]] if obs_ptr[j + 1] > obs_ptr[j] else -1 for j in range(P)], np.int32)
    return dict(kf_uid=kf_uid, kf_id=kf_id, kf_bad=(rng.random(K) < bad_kf).astype(np.uint8), kf_Tcw=Tcw, kf_intr=p.intr.astype(np.float32),
                kp_ptr=kp_ptr, kp_uv=kp_uv, kp_octave=kp_oct, inv_level_sigma2=table, kf_parent=None, loop_ptr=None, loop_kf=None,
                cov_ptr=None, cov_kf=None, cov_w=None, mp_uid=(1000000 * 4 + np.arange(P)).astype(np.int64), mp_id=mp_id,
                mp_bad=(rng.random(P) < bad_mp).astype(np.uint8), mp_pos=p.points.astype(np.float32), mp_ref=ref, obs_ptr=obs_ptr, obs_kf=okf,
                obs_idx=obs_idx, origin=0, map_id=map_id)


def flat_from_scene(sc, oracle, kf_rows, fixed_of_row, mp_rows_rule):
    """the flat BA problem a flattening rule yields: kf_rows = keyframe indices in row order; mp_rows_rule(j, obs) -> the observations to
    keep for point j (list of positions into its observation list) or None to leave the point out"""
    row_of = {k: r for r, k in enumerate(kf_rows)}
    poses = np.stack([oracle.pose_from_Tcw_f32(sc["kf_Tcw"][k]) for k in kf_rows])
    intr = sc["kf_intr"][kf_rows].astype(np.float64)
    pts, okf, omp, ouv, ow, mp_of_row = [], [], [], [], [], []
    for j in range(len(sc["mp_uid"])):
        obs = list(range(sc["obs_ptr"][j], sc["obs_ptr"][j + 1]))
        keep = mp_rows_rule(j, obs)
        if keep is None:
            continue
        r = len(pts); pts.append(sc["mp_pos"][j].astype(np.float64)); mp_of_row.append(j)
        for q in keep:
            k = int(sc["obs_kf"][q]); kp = sc["kp_ptr"][k] + sc["obs_idx"][q]
            okf.append(row_of[k]); omp.append(r); ouv.append(sc["kp_uv"][kp]); ow.append(sc["inv_level_sigma2"][sc["kp_octave"][kp]])
    p = synth.BAProblem(poses=poses, intr=intr, fixed=np.asarray(fixed_of_row, np.uint8), points=np.array(pts).reshape(-1, 3),
                        obs_kf=np.array(okf, np.int32), obs_mp=np.array(omp, np.int32), obs_uv=np.array(ouv, np.float32).reshape(-1, 2),
                        obs_w=np.array(ow, np.float32))
    return p, mp_of_row


def run_gba(sc, which, iterations, robust, loop, want_rc=0):
    """want_rc = -1: the callee is expected to throw (estd::infrastructure_ex); the wrapper reports that as -1"""
    keep = []
    S = c_scene(sc, keep)
    o, O = new_out(S.K, S.P)
    rc = lib().optw_gba(C.byref(S), int(which), int(iterations), int(robust), C.c_int64(loop[0]), C.c_int64(loop[1]), C.byref(O))
    assert rc == want_rc, rc
    return o


def run_gba_flip(sc, which, iterations, robust, loop, flip_kf):
    """As run_gba, but keyframe flip_kf turns bad between the solve and the write-back (device double's hook; shim library only)."""
    keep = []
    S = c_scene(sc, keep)
    o, O = new_out(S.K, S.P)
    rc = lib().optw_gba_flip(C.byref(S), int(which), int(iterations), int(robust), C.c_int64(loop[0]), C.c_int64(loop[1]), int(flip_kf), C.byref(O))
    assert rc == 0, rc
    return o


def run_gba_mirror(sc, iterations, robust, loop):
    """MapFusionGBA with a persistent map mirror registered for the map (shim library only): the problem comes from ccm_mirror_ba_problem"""
    keep = []
    S = c_scene(sc, keep)
    o, O = new_out(S.K, S.P)
    rc = lib().optw_gba_mirror(C.byref(S), int(iterations), int(robust), C.c_int64(loop[0]), C.c_int64(loop[1]), C.byref(O))
    assert rc == 0, rc
    return o


def run_gba_twice(sc, use_mirror, iterations):
    """two MapFusionGBA calls in a row (direct write-back); returns (map state, solver handles created: -1 if the library cannot tell)"""
    keep = []
    S = c_scene(sc, keep)
    o, O = new_out(S.K, S.P)
    n = C.c_int32(-2)
    rc = lib().optw_gba_twice(C.byref(S), int(use_mirror), int(iterations), C.byref(O), C.byref(n))
    assert rc == 0, rc
    return o, n.value


def run_essential_graph(sc, loop_kf, cur_kf, conn, fix_scale, loop_closure=False, corr=None, mp_corr_ref=None):
    """conn: {keyframe index: [keyframe indices]} = LoopConnections; corr = (kf indices, corrected (n,8), noncorrected (n,8))"""
    keep = []
    S = c_scene(sc, keep)
    o, O = new_out(S.K, S.P)
    ptr = np.zeros(S.K + 1, np.int32); flat = []
    for k in range(S.K):
        flat += sorted(conn.get(k, [])); ptr[k + 1] = len(flat)
    flat = np.asarray(flat, np.int32)
    if corr is None:
        ck, cc, cn = np.zeros(0, np.int32), np.zeros((0, 8)), np.zeros((0, 8))
    else:
        ck, cc, cn = np.ascontiguousarray(corr[0], np.int32), np.ascontiguousarray(corr[1], np.float64), np.ascontiguousarray(corr[2], np.float64)
    mref = None if mp_corr_ref is None else np.ascontiguousarray(mp_corr_ref, np.int32)
    rc = lib().optw_essential_graph(C.byref(S), int(loop_kf), int(cur_kf), _p(ptr), _p(flat), int(fix_scale), int(loop_closure), len(ck), _p(ck),
                                    _p(cc), _p(cn), _p(mref), C.byref(O))
    assert rc == 0
    return o


def run_local_ba(sc, kf_index, server=False):
    keep = []
    S = c_scene(sc, keep)
    o, O = new_out(S.K, S.P)
    assert lib().optw_local_ba(C.byref(S), int(kf_index), int(server), C.byref(O)) == 0
    return o


def run_pose_optimization(sc, kp_uv, kp_octave, mp_of_kp, Tcw32, intr):
    keep = []
    S = c_scene(sc, keep)
    uv = np.ascontiguousarray(kp_uv, np.float32); oc = np.ascontiguousarray(kp_octave, np.int32); mp = np.ascontiguousarray(mp_of_kp, np.int32)
    T = np.ascontiguousarray(Tcw32, np.float32); K4 = np.ascontiguousarray(intr, np.float32)
    Tout = np.zeros((4, 4), np.float32); outl = np.zeros(len(oc), np.uint8); nsp = C.c_int32(0)
    r = lib().optw_pose_optimization(C.byref(S), len(oc), _p(uv), _p(oc), _p(mp), _p(T), _p(K4), _p(Tout), _p(outl), C.byref(nsp))
    return r, Tout, outl, nsp.value


def run_optimize_sim3(sc, k1, k2, match1, S12, th2, fix_scale):
    keep = []
    S = c_scene(sc, keep)
    m = np.ascontiguousarray(match1, np.int32); s_in = np.ascontiguousarray(S12, np.float64); s_out = np.zeros(8); m_out = np.zeros(len(m), np.int32)
    r = lib().optw_optimize_sim3(C.byref(S), int(k1), int(k2), _p(m), _p(s_in), C.c_float(th2), int(fix_scale), _p(s_out), _p(m_out))
    return r, s_out, m_out
