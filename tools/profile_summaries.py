"""Turn the raw profiler outputs brought back from the GPU box into the tracked summaries under profiles/.

  python tools/profile_summaries.py launches <launches.csv> <bench.json> <out.md>
      ncu --metrics gpu__time_duration.sum launch list -> per-kernel table, with the CUDA-event shares bench.py measured
  python tools/profile_summaries.py full <out.md> <traffic.json> <workload> <report.ncu-rep>...
      ncu --set full reports -> key metrics per kernel + DRAM traffic per launch (read by bench.py for roofline.traffic)
"""
from __future__ import annotations

import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict

BENCH_KEYS = {  # kernel-name fragment -> key in bench.py's "kernels" table
    "k_pcg": "pcg", "k_schur": "schur", "k_scale": "scale", "k_linearize": "linearize", "k_backsub_points": "backsub",
    "k_pose_pass": "pose_pass", "k_residual": "residual", "k_finalize_S": "finalize",
}


def short(name: str) -> str:
    m = re.search(r"(k_[A-Za-z0-9_]+(<[^>]*>)?)", name)
    return m.group(1) if m else name[:40]


def launches(csv_path: str, bench_path: str, out_path: str) -> None:
    rows = [l for l in open(csv_path) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    tot: "OrderedDict[str, list]" = OrderedDict()
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        t = tot.setdefault(k, [0, 0.0])
        t[0] += 1
        t[1] += float(r["Metric Value"].replace(",", "")) * 1e-6
    bench = json.loads(open(bench_path).read().strip().splitlines()[-1])
    shares = {k: v["share"] for k, v in bench.get("kernels", {}).items()}
    lm_kernels = [k for k in tot if any(k.startswith(f) for f in BENCH_KEYS)]
    lm_total = sum(tot[k][1] for k in lm_kernels)
    with open(out_path, "w") as f:
        f.write("# Launch list under `ncu --metrics gpu__time_duration.sum --clock-control none` (%s, 1xB200) - raw: %s\n\n" % (
            bench["config"]["workload"], csv_path.split("/")[-1]))
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES with the CUDA-event shares bench.py\n"
                "reports (`kernels.*.share`), not absolutes.  ncu share = kernel total / total of the LM-loop kernels listed in\n"
                "bench.py's table; structure kernels (run once per handle) are listed without a share.\n\n")
        f.write("| kernel | launches | total ms | share of LM loop (ncu) | share (bench.py, CUDA events) |\n|---|---|---|---|---|\n")
        for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            key = next((v for frag, v in BENCH_KEYS.items() if k.startswith(frag)), None)
            f.write("| %s | %d | %.3f | %s | %s |\n" % (k, n, ms, "%.3f" % (ms / lm_total) if key else "",
                                                     "%.3f" % shares[key] if key in shares else ""))
    print("wrote", out_path)


METRICS = OrderedDict([
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of ncu peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2 %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("lts__t_sector_hit_rate.pct", "l2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "l1 hit %"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
])
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def full(out_path: str, traffic_path: str, workload: str, reports: list) -> None:
    traffic = {}
    with open(out_path, "w") as f:
        f.write("# ncu --set full summaries (%s; 1xB200; --clock-control none)\n\n" % workload)
        f.write("Reports: " + ", ".join("`profiles/%s`" % r.split("/")[-1] for r in reports) + "\n")
        for rep in reports:
            out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rows = list(csv.reader(io.StringIO(out)))
            hdr, units, vals = rows[0], rows[1], rows[2]
            col = {h: i for i, h in enumerate(hdr)}
            name = vals[col["Kernel Name"]]
            f.write("\n## %s\n\n" % name[:80])
            rd = wr = None
            for m, label in METRICS.items():
                if m not in col:
                    continue
                v, u = vals[col[m]], units[col[m]]
                f.write("- %s: %s %s\n" % (label, v, u))
                if m == "dram__bytes_read.sum":
                    rd = float(v.replace(",", "")) * UNIT_SCALE.get(u, 1.0)
                if m == "dram__bytes_write.sum":
                    wr = float(v.replace(",", "")) * UNIT_SCALE.get(u, 1.0)
            if rd is not None and wr is not None:
                f.write("- traffic_bytes (read + write, this launch): %.0f\n" % (rd + wr))
                key = next((v for frag, v in BENCH_KEYS.items() if frag in name), None)
                if key:
                    traffic[key] = rd + wr
    old = {}
    try:
        old = json.load(open(traffic_path))
    except (OSError, ValueError):
        pass
    old.setdefault(workload, {}).update(traffic)
    old["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch from one ncu --set full capture each "
                    "(profiles/*.ncu-rep), 1xB200")
    json.dump(old, open(traffic_path, "w"), indent=1)
    print("wrote", out_path, traffic_path)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(*sys.argv[2:5])
    elif sys.argv[1] == "full":
        full(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5:])
    else:
        raise SystemExit(__doc__)
