"""Developer probe: ms per launch of every Schur-product kernel variant (ccm_ba_debug_set_schur_mode 0..5) on one config, and one
Global BA with the default.  With --capture it only prepares the handle and launches the default variant a few times (for ncu -k)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccm_slam_b200 import api, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "cfg5"
capture = "--capture" in sys.argv
api.init(0)
t = time.time(); p = synth.make_config(name); print(f"[{name}] K={p.K} P={p.P} E={p.E} generated in {time.time() - t:.1f}s", flush=True)
h = api.BAHandle(p)
info = h.info()
if capture:
    print("capture run:", h.time_kernel(4, reps=2, lam=1e-3), "ms")
    sys.exit(0)
labels = {0: "gather", 1: "mma u8 cta128 (default)", 2: "mma u16 cta256", 3: "mma u4 cta256", 4: "mma u8 cta256", 5: "mma u8 cta512", 6: "mma u8 cta64", 7: "mma u16 cta128", 8: "mma u8 cta128 prefetch"}
out = {"config": name, "schur_products": int(info["schur_products"]), "upper_blocks": int(info["s_blocks_upper"]), "ms_per_launch": {}}
for mode in (1, 0, 2, 3, 4, 5, 6, 7, 8, 1):
    api._chk(api.lib().ccm_ba_debug_set_schur_mode(mode))
    ms = h.time_kernel(4, reps=5, lam=1e-3)
    out["ms_per_launch"].setdefault(labels[mode], []).append(round(ms, 4))
api._chk(api.lib().ccm_ba_debug_set_schur_mode(-1))
h.reset(); h.set_profile(True)
t = time.time(); r = h.optimize(iterations=20, want_state=False); wall = time.time() - t
st = h.kernel_stats()
tot = sum(v["total_ms"] for v in st.values())
out["global_ba_default"] = dict(wall_s=round(wall, 4), iters=int(r["iters_done"]), trials=int(r["trials_total"]), pcg_iters=int(r["pcg_iters_total"]),
                                kernels_ms={k: round(v["total_ms"], 3) for k, v in st.items() if v["launches"]}, all_kernels_ms=round(tot, 3),
                                lm_iters_per_s=round(r["iters_done"] / (tot * 1e-3), 3))
print("RESULT " + json.dumps(out))
