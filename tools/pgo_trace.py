"""Developer probe: LM traces of the Sim3 pose graph, device vs oracle (the a4 parity test accepts different iteration counts: which decision diverges?)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ccm_slam_b200 import api, synth
from oracle import pyoracle
api.init(0)
np.set_printoptions(linewidth=200, precision=12)
for K, fs in ((60, False), (200, False), (200, True)):
    p = synth.make_pgo(K=K, fix_scale=fs)
    ref = pyoracle.pgo_solve(p, iterations=20)
    got = api.pgo_solve(p, iterations=20)
    print(f"== K={K} fix_scale={fs}: iters oracle {ref['iters_done']} device {got['iters_done']}  chi2_final {ref['chi2_final']:.15g} / {got['chi2_final']:.15g}")
    print("oracle trace [it, lambda, chi2, rho, trials, lambda_after]:\n", ref["trace"][:ref["iters_done"], :6])
    print("device trace:\n", got["trace"][:got["iters_done"], :6])
    print("max |dsim3| =", np.abs(got["sim3"] - ref["sim3"]).max())
