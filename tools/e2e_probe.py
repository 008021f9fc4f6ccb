"""Break the one-shot ccm_ba_solve call into its phases (setup / optimize / download / rest) on one workload."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccm_slam_b200 import api, synth

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
api.init(0)
cfg = dict(synth.CONFIGS[name]); cfg.pop("kind", None)
p = synth.make_global_ba(**cfg)
arrs = [p.poses, p.intr, p.fixed, p.points, p.obs_kf, p.obs_mp, p.obs_uv, p.obs_w]
for a in arrs:
    api.host_register(a)
api.ba_solve(p, iterations=1, huber_delta=api.HUBER_GBA, want_edges=False)
for k in range(6):
    t0 = time.perf_counter()
    r = api.ba_solve(p, iterations=8, huber_delta=api.HUBER_GBA, want_edges=False)
    w = (time.perf_counter() - t0) * 1e3
    print("wall %.1f ms  setup %.1f  optimize %.1f (event %.1f)  download %.1f  pcg_its %d" % (
        w, r["t_setup_ms"], r["t_optimize_ms"], r["t_optimize_event_ms"], r["t_download_ms"], r["pcg_iters_total"]))
