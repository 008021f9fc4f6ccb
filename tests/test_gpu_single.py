"""GPU parity tests of the single-vertex optimisations (PoseOptimizationClient, OptimizeSim3): C ABI vs CPU oracle.
Inlier / outlier sets and inlier counts must be identical; estimates within 1e-4 relative (north_star), in practice ~1e-10 for the
analytic pose problem and ~1e-6 for the Sim3 problem whose numeric Jacobians (delta 1e-9) amplify rounding."""
import numpy as np
import pytest

from ccm_slam_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _dev():
    assert api.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    api.init(0)


def _close(a, b, tol):
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


POSE_CASES = [dict(n=300, seed=11, outlier_frac=0.15), dict(n=60, seed=12, outlier_frac=0.3), dict(n=1000, seed=13, outlier_frac=0.05),
              dict(n=12, seed=14, outlier_frac=0.0), dict(n=8, seed=4, outlier_frac=0.0), dict(n=2, seed=3), dict(n=2500, seed=15, outlier_frac=0.1)]


def test_pose_optimization_batch_matches_oracle(oracle):
    probs = [synth.make_pose_opt(**c) for c in POSE_CASES]
    got = api.pose_optimize(probs)                       # one launch, one CTA per frame
    for d, (T, out, nin) in zip(probs, got):
        rT, rout, rnin = oracle.pose_optimize(d["Tcw0"], d["Xw"], d["uv"], d["inv_sigma2"], d["intr"])
        assert nin == rnin and np.array_equal(out, rout)
        assert _close(T, rT, 1e-4) and _close(T, rT, 1e-8)
    assert got[5][2] == 0 and np.array_equal(got[5][0], probs[5]["Tcw0"])   # n < 3: untouched
    one = api.pose_optimize(probs[:1])[0]                # batch of one gives the same answer as inside a batch
    assert np.array_equal(one[0], got[0][0]) and np.array_equal(one[1], got[0][1])
    assert api.pose_optimize([]) == []


SIM3_CASES = [dict(n=120, seed=12), dict(n=120, seed=12, fix_scale=True), dict(n=40, seed=21), dict(n=12, seed=5, outlier_frac=0.6),
              dict(n=600, seed=22, outlier_frac=0.1)]


def test_sim3_optimization_batch_matches_oracle(oracle):
    probs = [synth.make_sim3_opt(**c) for c in SIM3_CASES]
    got = api.sim3_optimize(probs)
    for d, (S, inl, nin) in zip(probs, got):
        rS, rinl, rnin = oracle.sim3_optimize(d["S12_0"], d["P1c"], d["P2c"], d["uv1"], d["uv2"], d["w1"], d["w2"], d["K1"], d["K2"],
                                              d["th2"], d["fix_scale"])
        assert nin == rnin and np.array_equal(inl, rinl)
        assert _close(S, rS, 1e-4)
        if d["fix_scale"]:
            assert S[7] == d["S12_0"][7]
    assert api.sim3_optimize([]) == []
