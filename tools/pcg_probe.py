"""Developer probe: one Global BA (optimize(20)) per variant on one config, PCG iterations, kernel ms and the PCG phase cycles.
Usage: python tools/pcg_probe.py cfg5 "CCM_PCG_IMPL=1" "CCM_PCG_NC=256,CCM_PCG_REFRESH=8" ...   ("-" = defaults)
The environment switches are read when the handle is created, so every variant gets its own handle on the same problem."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from ccm_slam_b200 import api, synth  # noqa: E402

name = sys.argv[1]
variants = sys.argv[2:] or ["-"]
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
api.init(local)
if world > 1:   # under torchrun: landmark shards + row-distributed PCG, every rank runs every variant, rank 0 reports
    import torch
    import torch.distributed as dist
    dist.init_process_group(backend="gloo")
    uid = torch.from_numpy(api.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8))
    dist.broadcast(uid, src=0)
    api.comm_init(rank, world, uid.numpy())
t = time.time(); p = synth.make_config(name); print(f"[{name}] K={p.K} P={p.P} E={p.E} generated in {time.time() - t:.1f}s", flush=True)
os.environ["CCM_PCG_PROF"] = "1"
ref = None
for v in variants:
    keys = []
    if v != "-":
        for kv in v.split(","):
            k, val = kv.split("="); os.environ[k] = val; keys.append(k)
    try:
        h = api.BAHandle(p)
        h.optimize(iterations=20, want_state=False)          # warm
        h.reset(); h.set_profile(True)
        r = h.optimize(iterations=20, want_state=True)
        st = h.kernel_stats()
        tot = sum(x["total_ms"] for x in st.values())
        cyc = h.pcg_cycles() if hasattr(h, "pcg_cycles") else None
        if ref is None:
            ref = r
        dp = float(np.abs(r["poses"] - ref["poses"]).max()); dx = float(np.abs(r["points"] - ref["points"]).max())
        if rank == 0:
          print("RESULT " + json.dumps({"variant": v, "world": world, "iters": int(r["iters_done"]), "trials": int(r["trials_total"]), "pcg_iters": int(r["pcg_iters_total"]),
                                      "pcg_not_converged": int(r["pcg_not_converged"]), "chi2_final": r["chi2_final"],
                                      "kernels_ms": {k: round(x["total_ms"], 3) for k, x in st.items() if x["launches"]},
                                      "all_kernels_ms": round(tot, 3), "event_ms": round(r["t_optimize_event_ms"], 3),
                                      "lm_iters_per_s": round(r["iters_done"] / (r["t_optimize_event_ms"] * 1e-3), 3),
                                      "pcg_us_per_iter": round(1e3 * st["pcg"]["total_ms"] / max(r["pcg_iters_total"], 1), 2),
                                      "pcg_cycles": cyc, "max_abs_diff_vs_first": [dp, dx]}), flush=True)
        h.close()
    except Exception as e:  # keep going: the other variants still tell something
        if rank == 0:
          print("RESULT " + json.dumps({"variant": v, "world": world, "error": str(e)}), flush=True)
    for k in keys:
        os.environ.pop(k, None)
