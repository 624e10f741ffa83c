#!/usr/bin/env python
"""Prints the metrics we quote from an .ncu-rep (run where ncu is installed):
   python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--source N]"""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_bytes.sum", "lts__t_bytes.sum", "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
STALL = "smsp__average_warps_issue_stalled_"

def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("====", r[idx["Kernel Name"]][:90], "grid", r[idx["Grid Size"]], "block", r[idx["Block Size"]])
        for k in KEYS:
            if k in idx:
                print(f"  {k:70s} {r[idx[k]]:>16s} {units[idx[k]]}")
        st = [(h[len(STALL):].replace("_per_issue_active.ratio", ""), float(r[i])) for h, i in idx.items() if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i]]
        st.sort(key=lambda x: -x[1])
        print("  stalls per issue:", ", ".join(f"{a}={b:.2f}" for a, b in st[:8]))
    if "--source" in sys.argv:
        topn = int(sys.argv[sys.argv.index("--source") + 1])
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
        print(src[:200])

if __name__ == "__main__":
    main()
