"""Developer probe: the two Schur-product kernels (gather form vs one f64 mma.sync per product) on one config.
Prints ms per launch of each, the agreement of what they produce (bschur, the PCG step) and a full Global BA in both modes."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccm_slam_b200 import api, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
api.init(0)
t = time.time(); p = synth.make_config(name); print(f"[{name}] K={p.K} P={p.P} E={p.E} generated in {time.time() - t:.1f}s", flush=True)
t = time.time(); h = api.BAHandle(p); info = h.info(); print(f"create {time.time() - t:.2f}s {info}", flush=True)
out = {"config": name, "info": {k: int(v) for k, v in info.items()}}
lam = 1e-3
step = {}
for mode, label in ((0, "gather"), (1, "mma")):
    api._chk(api.lib().ccm_ba_debug_set_schur_mode(mode))
    ms = h.time_kernel(4, reps=5, lam=lam)
    d = h.debug_schur(lam)
    step[label] = d
    h.reset(); h.set_profile(True)
    t = time.time(); r = h.optimize(iterations=20, want_state=True); wall = time.time() - t
    st = h.kernel_stats()
    out[label] = dict(schur_ms_per_launch=ms, ba_wall_s=wall, iters=int(r["iters_done"]), trials=int(r["trials_total"]),
                      pcg_iters=int(r["pcg_iters_total"]), chi2_final=float(r["chi2_final"]),
                      schur_total_ms=st["schur"]["total_ms"], schur_launches=st["schur"]["launches"], pcg_total_ms=st["pcg"]["total_ms"],
                      all_kernels_ms=float(sum(v["total_ms"] for v in st.values())))
    step[label + "_poses"] = r["poses"]; step[label + "_points"] = r["points"]
    print(label, json.dumps(out[label]), flush=True)
    h.reset()
api._chk(api.lib().ccm_ba_debug_set_schur_mode(-1))
rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
out["agreement"] = dict(bschur=rel(step["mma"]["bschur"], step["gather"]["bschur"]), dx_pose=rel(step["mma"]["dx_pose"], step["gather"]["dx_pose"]),
                        dx_point=rel(step["mma"]["dx_point"], step["gather"]["dx_point"]),
                        final_poses=rel(step["mma_poses"], step["gather_poses"]), final_points=rel(step["mma_points"], step["gather_points"]))
out["speedup_schur"] = out["gather"]["schur_ms_per_launch"] / out["mma"]["schur_ms_per_launch"]
print("RESULT " + json.dumps(out))
