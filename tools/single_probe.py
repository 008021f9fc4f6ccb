"""Latency / throughput of the single-vertex optimisations: C ABI call (host buffers, copies inside) vs the CPU oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ccm_slam_b200 import api, synth
from oracle import pyoracle as orc

api.init(0); orc.lib()


def bench(fn, reps=20):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


for n in (300, 1000):
    probs = [synth.make_pose_opt(n=n, seed=100 + i, outlier_frac=0.1) for i in range(64)]
    d = probs[0]
    t_cpu = bench(lambda: orc.pose_optimize(d["Tcw0"], d["Xw"], d["uv"], d["inv_sigma2"], d["intr"]))
    t1 = bench(lambda: api.pose_optimize(probs[:1]))
    t64 = bench(lambda: api.pose_optimize(probs), reps=5)
    print("PoseOptimization n=%4d: oracle %.3f ms/frame | GPU batch 1: %.3f ms | batch 64: %.3f ms (%.4f ms/frame)" % (n, t_cpu, t1, t64, t64 / 64))
for n in (120, 600):
    probs = [synth.make_sim3_opt(n=n, seed=200 + i) for i in range(32)]
    d = probs[0]
    t_cpu = bench(lambda: orc.sim3_optimize(d["S12_0"], d["P1c"], d["P2c"], d["uv1"], d["uv2"], d["w1"], d["w2"], d["K1"], d["K2"], d["th2"], d["fix_scale"]), reps=5)
    t1 = bench(lambda: api.sim3_optimize(probs[:1]))
    t32 = bench(lambda: api.sim3_optimize(probs), reps=5)
    print("OptimizeSim3     n=%4d: oracle %.3f ms/pair  | GPU batch 1: %.3f ms | batch 32: %.3f ms (%.4f ms/pair)" % (n, t_cpu, t1, t32, t32 / 32))
