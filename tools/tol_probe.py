"""How far the GPU estimate moves from the CPU oracle (exact sparse LDL^T per trial) as the PCG tolerance is loosened."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ccm_slam_b200 import api, synth
from oracle import pyoracle

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 8
api.init(0)
pyoracle.lib(); orc = pyoracle
p = synth.make_config(name)
t0 = time.perf_counter()
if os.environ.get("TOL_REF", "oracle") == "oracle":
    ref = orc.ba_solve(p, iterations=iters, huber_delta=api.HUBER_GBA)
    print("reference = oracle, %.1f s, iters %d" % (time.perf_counter() - t0, ref["iters_done"]))
else:  # sizes the oracle cannot finish in seconds: reference = the GPU path with PCG run to 1e-13
    ref = api.ba_solve(p, iterations=iters, huber_delta=api.HUBER_GBA, pcg_tol=1e-13, pcg_max_iter=5000, want_edges=False)
    print("reference = GPU at pcg_tol 1e-13, pcg its %d, not converged %d" % (ref["pcg_iters_total"], ref["pcg_not_converged"]))


def rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


for tol in (1e-10, 1e-8, 1e-6, 1e-4):
    r = api.ba_solve(p, iterations=iters, huber_delta=api.HUBER_GBA, pcg_tol=tol, want_edges=False)
    print("tol %.0e: pcg its %5d  optimize %.1f ms  pose rel %.2e  point rel %.2e  chi2 rel %.2e  iters %d trials %d" % (
        tol, r["pcg_iters_total"], r["t_optimize_ms"], rel(r["poses"], ref["poses"]), rel(r["points"], ref["points"]),
        abs(r["trace"][r["iters_done"] - 1, 2] - ref["trace"][ref["iters_done"] - 1, 2]) / ref["trace"][ref["iters_done"] - 1, 2],
        r["iters_done"], r["trials_total"]))
