"""Oracle for the single-vertex optimisations (PoseOptimizationClient, OptimizeSim3) against independent numpy witnesses
(tests/witness.py: chain-rule Jacobians, scipy rotations, Sim3 exponential through scipy.linalg.expm)."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from tests import witness as W
from ccm_slam_b200 import synth


@pytest.mark.parametrize("n,seed,frac", [(300, 11, 0.15), (60, 12, 0.3), (1000, 13, 0.05), (12, 14, 0.0)])
def test_pose_optimization_vs_witness(oracle, n, seed, frac):
    d = synth.make_pose_opt(n=n, seed=seed, outlier_frac=frac)
    T, outlier, nin = oracle.pose_optimize(d["Tcw0"], d["Xw"], d["uv"], d["inv_sigma2"], d["intr"])
    (R, t), wout, wnin = W.pose_optimization(d)
    assert nin == wnin and np.array_equal(outlier.astype(bool), wout)
    assert np.abs(Rotation.from_quat(T[:4]).as_matrix() - R).max() < 1e-9 and np.abs(T[4:] - t).max() < 1e-9
    if frac > 0:
        assert 0 < outlier.sum() < n
    # the estimate moved towards the ground truth
    assert np.abs(T[4:] - d["Tcw_gt"][4:]).max() < np.abs(d["Tcw0"][4:] - d["Tcw_gt"][4:]).max()


def test_pose_optimization_edge_cases(oracle):
    d = synth.make_pose_opt(n=2, seed=3)
    T, outlier, nin = oracle.pose_optimize(d["Tcw0"], d["Xw"], d["uv"], d["inv_sigma2"], d["intr"])
    assert nin == 0 and np.array_equal(T, d["Tcw0"])                      # < 3 correspondences: return 0, pose untouched
    d = synth.make_pose_opt(n=8, seed=4, outlier_frac=0.0)                 # < 10 edges: a single round
    T, outlier, nin = oracle.pose_optimize(d["Tcw0"], d["Xw"], d["uv"], d["inv_sigma2"], d["intr"])
    (R, t), wout, wnin = W.pose_optimization(d)
    assert nin == wnin and np.abs(T[4:] - t).max() < 1e-9


@pytest.mark.parametrize("n,seed,fix_scale", [(120, 12, False), (120, 12, True), (40, 21, False)])
def test_sim3_optimization_vs_witness(oracle, n, seed, fix_scale):
    d = synth.make_sim3_opt(n=n, seed=seed, fix_scale=fix_scale)
    S, inl, nin = oracle.sim3_optimize(d["S12_0"], d["P1c"], d["P2c"], d["uv1"], d["uv2"], d["w1"], d["w2"], d["K1"], d["K2"], d["th2"], fix_scale)
    st, winl, wnin = W.sim3_optimization(d)
    assert nin == wnin > 10 and np.array_equal(inl.astype(bool), winl)
    A = S[7] * Rotation.from_quat(S[:4]).as_matrix()
    # numeric Jacobians (delta 1e-9) carry ~1e-7 relative noise into the LM path: compare at 1e-5
    assert np.abs(A - st[0]).max() < 1e-5 and np.abs(S[4:7] - st[1]).max() < 1e-5
    if fix_scale:
        assert S[7] == d["S12_0"][7]
    assert np.abs(S[4:7] - d["S12_gt"][4:7]).max() < 0.05


def test_sim3_optimization_too_few_inliers_returns_zero(oracle):
    d = synth.make_sim3_opt(n=12, seed=5, outlier_frac=0.6)
    S, inl, nin = oracle.sim3_optimize(d["S12_0"], d["P1c"], d["P2c"], d["uv1"], d["uv2"], d["w1"], d["w2"], d["K1"], d["K2"], d["th2"], False)
    st, winl, wnin = W.sim3_optimization(d)
    assert nin == wnin
    if nin == 0:
        assert np.array_equal(S, d["S12_0"])  # g2oS12 is only written when the second optimisation ran
