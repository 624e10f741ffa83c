#!/usr/bin/env python
"""Writes profiles/r02_ncu_linearize.json (what bench.py's roofline.traffic reads) from an `ncu --set full` report that
holds the linearize kernels at 20 M points:  python profiles/ncu_linearize_json.py gpurun_out/prof_c4_final_r02.ncu-rep"""
import csv, io, json, os, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, key, scale=None):
    v, u = float(r[ix[key]].replace(",", "")), units[ix[key]]
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
    return v * mult


out = {"source": f"{os.path.basename(rep)} (ncu --set full --clock-control none --import-source on, bench.py --workload c4; profiles/capture_r02_final.sh)",
       "points": 20000000, "algorithmic_bytes": 1280000000}
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    if "linearize_kernel" not in name:
        continue
    key = "linearize_kernel<fp32 maha, H+b+err>" if "<0, 1" in name else "linearize_kernel<fp32 maha, err only>"
    out[key] = {
        "gpu_time_us": round(val(r, "gpu__time_duration.sum"), 2),
        "dram_bytes_read": int(val(r, "dram__bytes_read.sum")),
        "dram_bytes_write": int(val(r, "dram__bytes_write.sum")),
        "dram_pct_of_peak": round(float(r[ix["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]), 1),
        "fp64_pipe_pct": round(float(r[ix["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]]), 1),
        "issue_active_pct": round(float(r[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]]), 1),
        "registers": int(float(r[ix["launch__registers_per_thread"]])),
    }
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "r02_ncu_linearize.json"), "w"), indent=2)
print(json.dumps(out, indent=1))
