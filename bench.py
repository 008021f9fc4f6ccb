#!/usr/bin/env python
"""bench.py — global-BA LM iterations / second on synthetic BA problems of the BASELINE.json shapes.

    python bench.py --gpus N --steps K --warmup W [--workload cfg5] [--impl reference]

A "step" is one Global-BA solve: optimizer.optimize(20) of Optimizer::MapFusionGBA (S/Optimizer.cpp:797) on the
workload (default cfg5 = synthetic 10k keyframes x 1M landmarks x 20M observations, the only BASELINE config defined
at 1/2/4/8 GPUs; it fits one B200).  value = LM iterations per second of the whole job with the problem resident in
HBM (handle API: reset estimate -> ccm_ba_optimize), max over ranks.  e2e = the same metric through the
reference-facing one-shot call ccm_ba_solve with HOST buffers (upload + structure build + LM + download inside the
timed region).  N>1: one process per GPU (torchrun), landmarks sharded, the rendezvous/timing plumbing uses
torch.distributed (gloo); the data path uses the library's own NCCL communicator.

--impl reference times the reference's CPU algorithm (the dependency-free oracle port; g2o itself cannot be built
here: no Eigen) on the host, single thread — the reference build is single-threaded by construction
(cslam/thirdparty/g2o/config.h:4) — on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from ccm_slam_b200 import synth  # noqa: E402

METRIC = "global-BA LM iters/sec"
UNIT = "LM iters/s"
LM_ITERS = 20  # Opt.GBAIterations, cslam/conf/config.yaml:129


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().strip().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def algorithmic_bytes(info, P_local, E_local):
    """Compulsory HBM bytes per launch of each kernel group (SURVEY.md §8(d) per-unit figures)."""
    K = info["K"]; Kf = info["K_free"]; nub = info["s_blocks_upper"]; nnzb = info["s_blocks_full"]; npr = info["schur_products"]
    E, P = E_local, P_local
    return {
        "linearize": E * (20 + 144) + P * (24 + 72) + K * 56,          # B_lin without the Hpp write (that is pose_pass)
        "pose_pass": E * 16 + P * 24 + K * 56 + Kf * 336,              # packed (u, v, w, lm) stream + point + Hpp/bp write
        "scale": E * (144 + 144 + 4) + P * (48 + 24 + 24),             # W read, Z write, Hll/bl read, g write
        "schur": E * 144 + npr * 8 + nub * 288 + Kf * 48,              # Z once, product lists, S upper blocks + bschur write
        "finalize": nub * 288 + nnzb * 288 + Kf * (336 + 288 + 48),
        "pcg": None,                                                    # iterations * (nnzb*288 + ~10 vectors): filled in below
        "backsub": E * (144 + 4) + P * (72 + 24 + 24) + K * 112,
        "residual": E * 20 + P * 24 + K * 56,
        "allreduce": nub * 288 + Kf * 48,
    }


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU algorithm (oracle port) on a bounded sample of the workload, rank 0 only."""
    if rank != 0:
        return
    from oracle import pyoracle
    cfg = dict(synth.CONFIGS[args.workload])
    scale = 1
    sample = f"{args.workload} full size"
    if cfg.get("K", 0) >= 5000:  # cfg5: 1/10 of the trajectory, same band structure and density
        scale = 10
        cfg["K"] //= scale; cfg["P"] //= scale
        sample = (f"{args.workload} at 1/{scale} trajectory length (K={cfg['K']}, P={cfg['P']}, same 20 obs/landmark, same band), "
                  f"time scaled x{scale} (per-iteration cost of the banded problem is linear in its length)")
    kind = cfg.pop("kind")
    p = synth.make_global_ba(name=args.workload, **cfg) if kind == "global" else synth.make_local_ba(name=args.workload, **cfg)
    its = 2 if scale > 1 else LM_ITERS
    steps = max(1, min(args.steps, 2))
    t_tot, it_tot = 0.0, 0
    for _ in range(min(args.warmup, 1)):
        pyoracle.ba_solve(p, iterations=1, huber_delta=float(np.float32(np.sqrt(5.99))))
    for _ in range(steps):
        t0 = time.perf_counter()
        r = pyoracle.ba_solve(p, iterations=its, huber_delta=float(np.float32(np.sqrt(5.99))))
        t_tot += time.perf_counter() - t0
        it_tot += r["iters_done"]
    value = it_tot / (t_tot * scale)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * t_tot * scale / steps * (LM_ITERS / its), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "K": synth.CONFIGS[args.workload].get("K"), "P": synth.CONFIGS[args.workload].get("P"),
                       "lm_iterations": LM_ITERS, "solver": "direct sparse LDL^T (as g2o LinearSolverEigen)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                             "note": "oracle port of the g2o path; g2o itself is not buildable here (no Eigen); reference build is single-threaded"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(args):
    from oracle import pyoracle
    cfg = dict(synth.CONFIGS[args.workload])
    scale = 1
    if cfg.get("K", 0) >= 5000:
        scale = 10
        cfg["K"] //= scale; cfg["P"] //= scale
    kind = cfg.pop("kind")
    p = synth.make_global_ba(**cfg) if kind == "global" else synth.make_local_ba(**cfg)
    its = 2 if scale > 1 else min(LM_ITERS, 10)
    t0 = time.perf_counter()
    r = pyoracle.ba_solve(p, iterations=its, huber_delta=float(np.float32(np.sqrt(5.99))))
    dt = time.perf_counter() - t0
    return {"value": r["iters_done"] / (dt * scale), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": (f"{args.workload} at 1/{scale} trajectory length (K={p.K}, P={p.P}, E={p.E}), {its} LM iterations, single thread, "
                       f"{dt:.1f} s of CPU work; time scaled x{scale}" if scale > 1 else
                       f"{args.workload} full size, {its} LM iterations, single thread, {dt:.1f} s of CPU work"),
            "breakdown_s": r["timing"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=sorted(synth.CONFIGS))
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    from ccm_slam_b200 import api
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        dist.init_process_group(backend="gloo")
    api.init(local_rank)
    if world > 1:
        import torch
        uid = torch.from_numpy(api.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8))
        dist.broadcast(uid, src=0)
        api.comm_init(rank, world, uid.numpy())

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def sum_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    p = synth.make_config(args.workload)
    delta = api.HUBER_GBA
    h = api.BAHandle(p)
    info = h.info()
    small = info["device_bytes"] < (200 << 20)  # inputs not larger than L2 -> flush between iterations

    def one_step():
        h.reset()
        if small:
            api.l2_flush()
        return h.optimize(iterations=LM_ITERS, huber_delta=delta, want_state=False)

    for _ in range(max(args.warmup, 3)):
        one_step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    h.set_profile(True)
    launches0 = api.kernel_launches()
    t_dev_ms, it_tot, tr_tot, pcg_tot, pcg_nc = 0.0, 0, 0, 0, 0
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        barrier()
        r = one_step()
        t_dev_ms += max_over_ranks(r["t_optimize_event_ms"])  # CUDA events on the launching stream, max over ranks
        it_tot += r["iters_done"]; tr_tot += r["trials_total"]; pcg_tot += r["pcg_iters_total"]; pcg_nc += r["pcg_not_converged"]
    barrier()
    wall = time.perf_counter() - wall0
    launches = api.kernel_launches() - launches0
    kstats = h.kernel_stats()
    h.set_profile(False)
    clocks = sampler.stop() if sampler else None
    value = it_tot / (t_dev_ms * 1e-3)

    # ---- end-to-end through the one-shot C ABI call with host buffers (pinned), copies inside the timed region
    arrs = [p.poses, p.intr, p.fixed, p.points, p.obs_kf, p.obs_mp, p.obs_uv, p.obs_w]
    h2d = int(sum(a.nbytes for a in arrs)); d2h = int(p.poses.nbytes + p.points.nbytes)
    h.close()
    pinned = []
    for a in arrs:
        try:
            api.host_register(a); pinned.append(a)
        except api.CCMError:
            pass
    api.ba_solve(p, iterations=LM_ITERS, huber_delta=delta, want_edges=False)  # warm: every code path of the timed call
    barrier()
    # wall-clock around the public call: a fresh box stalls the HOST now and then (lazily paged image, first-touch of driver
    # pages) for hundreds of ms, which has nothing to do with the path -> every step is listed, the MEDIAN step is reported
    e2e_steps_ms, e2e_it, setup_ms = [], 0, 0.0
    for _ in range(args.e2e_steps):
        barrier()
        t0 = time.perf_counter()
        r = api.ba_solve(p, iterations=LM_ITERS, huber_delta=delta, want_edges=False)
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e_steps_ms.append(dt * 1e3)
        e2e_it += r["iters_done"]; setup_ms += r["t_setup_ms"]
        if rank == 0:
            print("[bench] e2e step: wall %.1f ms (setup %.1f, optimize %.1f, download %.1f, pcg its %d)" % (
                dt * 1e3, r["t_setup_ms"], r["t_optimize_ms"], r["t_download_ms"], r["pcg_iters_total"]), file=sys.stderr)
    for a in pinned:
        api.host_unregister(a)
    e2e_ms = statistics.median(e2e_steps_ms)
    e2e_val = (e2e_it / args.e2e_steps) / (e2e_ms * 1e-3)
    launches_all = sum_over_ranks(float(launches))

    if rank != 0:
        return
    hbm_peak, peak_src = peaks()
    nlaunch = {k: max(v["launches"], 1) for k, v in kstats.items()}
    ab = algorithmic_bytes(info, info["P_local"], info["E_local"])
    ab["pcg"] = (pcg_tot / max(nlaunch["pcg"], 1)) * (info["s_blocks_full"] * 288 + info["K_free"] * 48 * 12)
    step_ms = sum(v["total_ms"] for v in kstats.values())
    kernels = {}
    for k, v in kstats.items():
        if v["launches"] == 0:
            continue
        avg_ms = v["total_ms"] / v["launches"]
        gbs = ab[k] / (avg_ms * 1e-3) / 1e9 if ab.get(k) else None
        kernels[k] = {"launches": v["launches"], "avg_ms": avg_ms, "share": v["total_ms"] / step_ms,
                      "alg_bytes_per_launch": ab.get(k), "achieved_gbs": gbs, "frac_of_hbm_peak": (gbs / hbm_peak if gbs else None)}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    dom = max((k for k in kernels if k != "allreduce"), key=lambda k: kernels[k]["share"])
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.workload, {}).get(dom)
    roof = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
            "frac": kernels[dom]["frac_of_hbm_peak"], "traffic": traffic, "peak_source": peak_src,
            "share_of_step": kernels[dom]["share"],
            "named_target_kernel": {"kernel": "linearize+pose_pass (K2 of SURVEY 8(d), B_lin = E*164 + P*96 + K*392)",
                                    "achieved": (ab["linearize"] + info["K_free"] * 336) / ((kernels["linearize"]["avg_ms"] + kernels["pose_pass"]["avg_ms"]) * 1e-3) / 1e9,
                                    "linearize_alone_gbs": kernels["linearize"]["achieved_gbs"]}}
    roof["named_target_kernel"]["frac"] = roof["named_target_kernel"]["achieved"] / hbm_peak
    cpu = None if args.no_cpu_baseline else cpu_baseline(args)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": t_dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "K": p.K, "P": p.P, "E": p.E, "lm_iterations_max": LM_ITERS,
                       "lm_iterations_done_per_step": it_tot / args.steps, "trials_per_step": tr_tot / args.steps,
                       "pcg_iters_per_step": pcg_tot / args.steps, "pcg_not_converged": pcg_nc, "huber": "sqrt(5.99)",
                       "l2": "flushed between steps" if small else "inputs > L2 (W+Z+product lists are GBs)",
                       "parallelism": f"landmark-shard x{world}"},
            "trials_per_s": tr_tot / (t_dev_ms * 1e-3), "wall_s_timed_region": wall,
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "stat": "median of the per-step wall times", "step_ms": e2e_steps_ms,
                    "setup_ms_per_step": setup_ms / args.e2e_steps, "steps": args.e2e_steps,
                    "call": "ccm_ba_solve (host buffers, pinned)"},
            "gpu_launches": int(launches_all),
            "roofline": roof, "kernels": kernels, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
