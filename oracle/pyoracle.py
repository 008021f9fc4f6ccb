"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
TRACE_COLS = 6


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return so


class _BAProblem(C.Structure):
    _fields_ = [("K", C.c_int32), ("P", C.c_int32), ("E", C.c_int32),
                ("poses", C.c_void_p), ("intr", C.c_void_p), ("fixed", C.c_void_p), ("points", C.c_void_p),
                ("obs_kf", C.c_void_p), ("obs_mp", C.c_void_p), ("obs_uv", C.c_void_p), ("obs_w", C.c_void_p),
                ("edge_flags", C.c_void_p)]


class _BAOptions(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("robust", C.c_int32), ("huber_delta", C.c_double),
                ("lambda_init", C.c_double), ("max_trials", C.c_int32), ("stop", C.c_void_p)]


class _BAResult(C.Structure):
    _fields_ = [("poses", C.c_void_p), ("points", C.c_void_p), ("chi2", C.c_void_p), ("depth_pos", C.c_void_p),
                ("trace", C.c_void_p), ("trace_cap", C.c_int32), ("trace_len", C.c_int32),
                ("iters_done", C.c_int32), ("trials_total", C.c_int32),
                ("chi2_initial", C.c_double), ("chi2_final", C.c_double), ("lambda_final", C.c_double),
                ("t_build_s", C.c_double), ("t_schur_s", C.c_double), ("t_solve_s", C.c_double),
                ("t_resid_s", C.c_double), ("t_total_s", C.c_double), ("t_structure_s", C.c_double)]


class _PGOProblem(C.Structure):
    _fields_ = [("K", C.c_int32), ("E", C.c_int32), ("sim3", C.c_void_p), ("fixed", C.c_void_p),
                ("edge_i", C.c_void_p), ("edge_j", C.c_void_p), ("meas", C.c_void_p), ("fix_scale", C.c_int32)]


class _PGOResult(C.Structure):
    _fields_ = [("sim3", C.c_void_p), ("trace", C.c_void_p), ("trace_cap", C.c_int32), ("trace_len", C.c_int32),
                ("iters_done", C.c_int32), ("chi2_initial", C.c_double), ("chi2_final", C.c_double),
                ("lambda_final", C.c_double), ("t_total_s", C.c_double)]


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_ba_linearize.restype = C.c_double
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _ba_struct(p, keep):
    arrs = dict(poses=np.ascontiguousarray(p.poses, np.float64), intr=np.ascontiguousarray(p.intr, np.float64),
                fixed=np.ascontiguousarray(p.fixed, np.uint8), points=np.ascontiguousarray(p.points, np.float64),
                obs_kf=np.ascontiguousarray(p.obs_kf, np.int32), obs_mp=np.ascontiguousarray(p.obs_mp, np.int32),
                obs_uv=np.ascontiguousarray(p.obs_uv, np.float32), obs_w=np.ascontiguousarray(p.obs_w, np.float32),
                edge_flags=None if p.edge_flags is None else np.ascontiguousarray(p.edge_flags, np.uint8))
    keep.append(arrs)
    return _BAProblem(p.K, p.P, p.E, *[_p(arrs[k]) for k in
                                      ("poses", "intr", "fixed", "points", "obs_kf", "obs_mp", "obs_uv", "obs_w", "edge_flags")])


def ba_solve(p, iterations=20, robust=True, huber_delta=np.sqrt(5.99), lambda_init=-1.0, max_trials=10,
             chi2_in=None, stop=None):
    keep = []
    prob = _ba_struct(p, keep)
    poses = np.empty((p.K, 7)); points = np.empty((p.P, 3))
    chi2 = np.zeros(p.E) if chi2_in is None else np.array(chi2_in, np.float64)
    depth = np.zeros(p.E, np.uint8)
    trace = np.zeros((max(iterations, 1), TRACE_COLS))
    opt = _BAOptions(iterations, int(robust), float(huber_delta), float(lambda_init), max_trials, _p(stop))
    res = _BAResult(_p(poses), _p(points), _p(chi2), _p(depth), _p(trace), trace.shape[0])
    rc = lib().orc_ba_solve(C.byref(prob), C.byref(opt), C.byref(res))
    assert rc == 0
    return dict(poses=poses, points=points, chi2=chi2, depth_pos=depth, trace=trace[:res.trace_len],
                iters_done=res.iters_done, trials_total=res.trials_total, chi2_initial=res.chi2_initial,
                chi2_final=res.chi2_final, lambda_final=res.lambda_final,
                timing=dict(build=res.t_build_s, schur=res.t_schur_s, solve=res.t_solve_s, resid=res.t_resid_s,
                            structure=res.t_structure_s, total=res.t_total_s))


def ba_linearize(p, robust=True, huber_delta=np.sqrt(5.99)):
    keep = []
    prob = _ba_struct(p, keep)
    err = np.empty((p.E, 2)); Jp = np.empty((p.E, 2, 6)); Jl = np.empty((p.E, 2, 3))
    rho1 = np.empty(p.E); chi2 = np.empty(p.E)
    tot = lib().orc_ba_linearize(C.byref(prob), int(robust), C.c_double(huber_delta), _p(err), _p(Jp), _p(Jl), _p(rho1), _p(chi2))
    return dict(err=err, Jpose=Jp, Jpoint=Jl, rho1=rho1, chi2=chi2, chi2_robust_sum=tot)


def ba_build(p, robust=True, huber_delta=np.sqrt(5.99)):
    keep = []
    prob = _ba_struct(p, keep)
    Hpp = np.empty((p.K, 6, 6)); bp = np.empty((p.K, 6)); Hll = np.empty((p.P, 3, 3)); bl = np.empty((p.P, 3))
    W = np.empty((p.E, 6, 3))
    lib().orc_ba_build(C.byref(prob), int(robust), C.c_double(huber_delta), _p(Hpp), _p(bp), _p(Hll), _p(bl), _p(W))
    return dict(Hpp=Hpp, bp=bp, Hll=Hll, bl=bl, W=W)


def ba_schur_solve(p, lam, robust=True, huber_delta=np.sqrt(5.99), dense=False):
    keep = []
    prob = _ba_struct(p, keep)
    dxp = np.empty((p.K, 6)); dxl = np.empty((p.P, 3))
    S = np.empty((6 * p.K, 6 * p.K)) if dense else None
    bs = np.empty(6 * p.K) if dense else None
    rc = lib().orc_ba_schur_solve(C.byref(prob), int(robust), C.c_double(huber_delta), C.c_double(lam), _p(dxp), _p(dxl), _p(S), _p(bs))
    return dict(rc=rc, dx_pose=dxp, dx_point=dxl, S=S, bschur=bs)


def _vec(fn, inp, nout):
    a = np.ascontiguousarray(inp, np.float64)
    o = np.empty(nout)
    fn(_p(a), _p(o))
    return o


def se3_exp(u): return _vec(lib().orc_se3_exp, u, 7)
def sim3_exp(u): return _vec(lib().orc_sim3_exp, u, 8)
def sim3_log(s): return _vec(lib().orc_sim3_log, s, 7)
def sim3_inv(s): return _vec(lib().orc_sim3_inv, s, 8)


def se3_mul(a, b):
    a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64); o = np.empty(7)
    lib().orc_se3_mul(_p(a), _p(b), _p(o)); return o


def sim3_mul(a, b):
    a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64); o = np.empty(8)
    lib().orc_sim3_mul(_p(a), _p(b), _p(o)); return o


def se3_map(qt, x):
    qt = np.ascontiguousarray(qt, np.float64); x = np.ascontiguousarray(x, np.float64); o = np.empty(3)
    lib().orc_se3_map(_p(qt), _p(x), _p(o)); return o


def pose_from_Tcw_f32(T):
    T = np.ascontiguousarray(T, np.float32); o = np.empty(7)
    lib().orc_pose_from_Tcw_f32(_p(T), _p(o)); return o


def pose_to_Tcw_f32(qt):
    qt = np.ascontiguousarray(qt, np.float64); o = np.empty((4, 4), np.float32)
    lib().orc_pose_to_Tcw_f32(_p(qt), _p(o)); return o


def huber(e, delta):
    o = np.empty(3)
    lib().orc_huber(C.c_double(e), C.c_double(delta), _p(o)); return o


def pgo_edge_error(meas, si, sj):
    m = np.ascontiguousarray(meas, np.float64); a = np.ascontiguousarray(si, np.float64); b = np.ascontiguousarray(sj, np.float64)
    o = np.empty(7)
    lib().orc_pgo_edge_error(_p(m), _p(a), _p(b), _p(o)); return o


def pgo_solve(p, iterations=20, lambda_init=1e-16, analytic_jac=False, stop=None):
    arrs = dict(sim3=np.ascontiguousarray(p.sim3, np.float64), fixed=np.ascontiguousarray(p.fixed, np.uint8),
                ei=np.ascontiguousarray(p.edge_i, np.int32), ej=np.ascontiguousarray(p.edge_j, np.int32),
                meas=np.ascontiguousarray(p.meas, np.float64))
    K, E = arrs["sim3"].shape[0], arrs["ei"].shape[0]
    prob = _PGOProblem(K, E, _p(arrs["sim3"]), _p(arrs["fixed"]), _p(arrs["ei"]), _p(arrs["ej"]), _p(arrs["meas"]), int(p.fix_scale))
    out = np.empty((K, 8)); trace = np.zeros((max(iterations, 1), TRACE_COLS))
    res = _PGOResult(_p(out), _p(trace), trace.shape[0])
    rc = lib().orc_pgo_solve(C.byref(prob), iterations, C.c_double(lambda_init), int(analytic_jac), _p(stop), C.byref(res))
    assert rc == 0
    return dict(sim3=out, trace=trace[:res.trace_len], iters_done=res.iters_done, chi2_initial=res.chi2_initial,
                chi2_final=res.chi2_final, lambda_final=res.lambda_final, t_total=res.t_total_s)
