"""Developer probe: the tiled Schur schedule (CCM_SCHUR=9, tile edge CCM_SCHUR_TILE) and the sorted product lists against the untiled
prefetch kernel (mode 8): ms per launch on one config and the final state of one Global BA per variant (must agree to rounding)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from ccm_slam_b200 import api, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
api.init(0)
t = time.time(); p = synth.make_config(name); print(f"[{name}] K={p.K} P={p.P} E={p.E} generated in {time.time() - t:.1f}s", flush=True)
ref = None
ALL = [("untiled prefetch, unsorted lists", {"CCM_SCHUR_SORT": "0"}, 8), ("untiled prefetch, sorted lists", {"CCM_SCHUR_SORT": "1"}, 8),
       ("tiled 2x2", {"CCM_SCHUR_TILE": "2"}, 9), ("tiled 3x3", {"CCM_SCHUR_TILE": "3"}, 9), ("tiled 4x4", {"CCM_SCHUR_TILE": "4"}, 9),
       ("tiled 4x4, unsorted lists", {"CCM_SCHUR_TILE": "4", "CCM_SCHUR_SORT": "0"}, 9), ("row-synchronous", {}, 10),
       ("vectorised entries u8", {}, 11), ("vectorised entries u16", {}, 12), ("vectorised entries u8 + predicated padding lanes", {}, 13),
       ("vectorised entries u8 + wide row loads", {}, 14),
       ("entries through shared memory (u8)", {}, 15), ("baseline: the default list kernel (mode 11)", {}, 11),
       ("grouped lists (4 blocks of a row per warp), entries by shuffle", {}, 16),
       ("grouped lists (4 blocks of a row per warp), entries through shared memory", {}, 17)]
# (the padded layouts of the rows of Z measured in profiles/r2/zlayout_cfg5.log were a switch of commit f23ac9d, removed afterwards)
want = sys.argv[2:]   # optional: substrings of the variant labels to run
for label, env, mode in [v for v in ALL if not want or any(w in v[0] for w in want)]:
    for k, v in env.items():
        os.environ[k] = v
    try:
        api._chk(api.lib().ccm_ba_debug_set_schur_mode(mode))     # before the handle: the tile schedule is built at create time for mode 9
        t0 = time.time(); h = api.BAHandle(p); t_create = time.time() - t0
        ms = [round(h.time_kernel(4, reps=5, lam=1e-3), 4) for _ in range(2)]
        ms_scale = round(h.time_kernel(3, reps=5, lam=1e-3), 4)
        ms_backsub = round(h.time_kernel(5, reps=5, lam=1e-3), 4)
        h.reset(); h.set_profile(True)
        r = h.optimize(iterations=20, want_state=True)
        st = h.kernel_stats()
        if ref is None:
            ref = r
        print("RESULT " + json.dumps({"variant": label, "schur_ms_per_launch": ms, "scale_ms": ms_scale, "backsub_ms": ms_backsub, "create_s": round(t_create, 3), "iters": int(r["iters_done"]),
                                      "pcg_iters": int(r["pcg_iters_total"]), "schur_ms_in_gba": round(st["schur"]["total_ms"], 3),
                                      "event_ms": round(r["t_optimize_event_ms"], 3),
                                      "max_abs_diff_vs_first": [float(np.abs(r["poses"] - ref["poses"]).max()), float(np.abs(r["points"] - ref["points"]).max())]}), flush=True)
        h.close()
    except Exception as e:
        print("RESULT " + json.dumps({"variant": label, "error": str(e)}), flush=True)
    api._chk(api.lib().ccm_ba_debug_set_schur_mode(-1))
    for k in env:
        os.environ.pop(k, None)
