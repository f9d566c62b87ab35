#!/usr/bin/env python
"""Turns the ncu captures of scripts/profile_gpu.sh (gpurun_out/) into the tracked summaries under profiles/:
   profiles/ncu_full_summary_<round>.txt   per-kernel metrics of one step (ncu --set full)
   profiles/launches_<round>.csv           launch list with durations + the per-kernel share of a step
   profiles/k1_traffic_<round>.json        DRAM bytes of one K1 launch (bench.py's roofline.traffic)
Usage: python scripts/summarize_ncu.py r1"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r1"
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
           "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_fp64.sum"]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def pdq_summary():
    """profiles/ncu_pdq_summary_<round>.txt from gpurun_out/prof_pdq_<round>.ncu-rep (scripts/profile_gpu.sh <round> pdq)."""
    rep = os.path.join(GO, f"prof_pdq_{R}.ncu-rep")
    if not os.path.exists(rep):
        return
    hdr, units, rows = raw_page(rep)
    name_i = hdr.index("Kernel Name")
    lines = [f"# ncu --set full --clock-control none, round {R[1:]}, B200: PDQ kernels (csrc/kp_pdq.cu), workload of scripts/pdq_bench.py",
             "# 16 images of 720x1280, 60 detections + 20 ground-truth boxes each; dense maps: 256 detections; first launch of each kernel",
             ""]
    seen = set()
    for r in rows:
        k = r[name_i].split("(")[0]
        if k in seen:
            continue
        seen.add(k)
        lines.append(f"## {r[name_i]}")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                lines.append(f"{m:80s} {r[i]:>16s} {units[i]}")
        lines.append("")
    with open(os.path.join(OUT, f"ncu_pdq_summary_{R}.txt"), "w") as f:
        f.write("\n".join(lines))
    print("\n".join(l for l in lines if l.startswith("##") or "time_duration" in l))


def main():
    pdq_summary()
    if len(sys.argv) > 2 and sys.argv[2] == "pdq":
        return
    rep = os.path.join(GO, f"prof_all_{R}.ncu-rep")
    hdr, units, rows = raw_page(rep)
    name_i = hdr.index("Kernel Name")
    lines = [f"# ncu --set full --clock-control none, round {R[1:]}, B200, workload bdd_covar_b32_k11 (B=32, N=10, A=172980, K=11)",
             "# command: see scripts/profile_gpu.sh (one step at a time: --pipeline 1); per-launch values, one launch of each kernel",
             ""]
    k1 = None
    for r in rows:
        lines.append(f"## {r[name_i]}")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                lines.append(f"{m:80s} {r[i]:>16s} {units[i]}")
        lines.append("")
        if "k1_moments" in r[name_i]:
            def val(m):
                i = hdr.index(m); v = float(r[i].replace(",", "")); u = units[i].lower()
                return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1}.get(u, 1)
            k1 = dict(kernel=r[name_i], workload="bdd_covar_b32_k11",
                      dram_bytes_per_launch=val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                      gpu_time_us=float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")),
                      source=f"profiles/ncu_full_summary_{R}.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)")
    with open(os.path.join(OUT, f"ncu_full_summary_{R}.txt"), "w") as f:
        f.write("\n".join(lines))
    if k1:
        with open(os.path.join(OUT, f"k1_traffic_{R}.json"), "w") as f:
            json.dump(k1, f, indent=1)
    # launch list: keep the csv, append the share of each kernel
    src = os.path.join(GO, f"launches_{R}.csv")
    text = [l for l in open(src) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(text))))
    tot, per = 0.0, {}
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r.get("Metric Unit", "us"), 1.0)
        v *= scale
        k = r["Kernel Name"].split("(")[0]
        if not any(t in k for t in ("k1_", "k2_", "k3_", "k4_", "scan_tiles", "prefilter", "val_")):
            continue
        per.setdefault(k, [0, 0.0]); per[k][0] += 1; per[k][1] += v; tot += v
    with open(os.path.join(OUT, f"launches_{R}.csv"), "w") as f:
        f.write("".join(text))
    with open(os.path.join(OUT, f"launch_share_{R}.txt"), "w") as f:
        f.write(f"# share of the stage kernels in the ncu launch list (profiles/launches_{R}.csv; cold-cache, serialised launches)\n")
        for k, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:60s} launches={n:4d} total_us={t:10.1f} mean_us={t / n:8.1f} share={t / tot:6.3f}\n")
    print(open(os.path.join(OUT, f"launch_share_{R}.txt")).read())


if __name__ == "__main__":
    main()
