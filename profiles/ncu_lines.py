#!/usr/bin/env python
"""Per-source-line summary of an ncu capture taken with --import-source on (kernels compiled with -lineinfo):
warp-instructions and stall samples per file and for the top lines, with the dominant stall reasons.

    python profiles/ncu_lines.py gpurun_out/prof.ncu-rep [--top 30] [--kernel substring]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    lines, files = [], {}
    cur, hdr = None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            continue
        if hdr is None or cur is None or r[0] in ("Function Name", "Kernel Name") or not r[0].strip().isdigit():
            continue
        try:
            inst = int(r[hdr["Instructions Executed"]])
            tinst = int(r[hdr["Thread Instructions Executed"]])
            samples = int(r[hdr["# Samples"]])
        except (ValueError, KeyError, IndexError):
            continue
        stalls = {}
        for h, i in hdr.items():
            if h.startswith("stall_") and "(Not Issued)" not in h and i < len(r):
                try:
                    stalls[h[6:]] = int(r[i])
                except ValueError:
                    pass
        lines.append((cur, int(r[0]), r[1].strip(), inst, tinst, samples, stalls))
        f = files.setdefault(cur, [0, 0])
        f[0] += inst
        f[1] += samples
    tot_i = sum(v[0] for v in files.values()) or 1
    tot_s = sum(v[1] for v in files.values()) or 1
    print(f"warp-instructions by source file (total {tot_i}), stall samples (total {tot_s}):")
    for f, (i, s) in sorted(files.items(), key=lambda kv: -kv[1][0]):
        print(f"  {f:28s} {i:10d} {100.0 * i / tot_i:5.1f} %   samples {s:7d} {100.0 * s / tot_s:5.1f} %")
    agg = {}
    for _, _, _, _, _, _, st in lines:
        for k, v in st.items():
            agg[k] = agg.get(k, 0) + v
    print("stall samples by reason:", ", ".join(f"{k}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for title, key in (("top source lines by warp-instructions", lambda l: -l[3]), ("top source lines by stall samples", lambda l: -l[5])):
        print(f"\n{title} (instr, share, active threads/instr, samples, top stalls):")
        for f, ln, src, inst, tinst, samples, st in sorted(lines, key=key)[:top]:
            ts = ", ".join(f"{k}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
            print(f"  {f}:{ln:<4d} {inst:9d} {100.0 * inst / tot_i:5.1f}% thr {tinst / max(1, inst):5.1f} smp {samples:6d} {100.0 * samples / tot_s:4.1f}% [{ts}] | {src[:90]}")


if __name__ == "__main__":
    main()
