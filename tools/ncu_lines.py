"""Aggregate the warp-stall samples and executed instructions of an ncu report per CUDA source
line (innermost inlined frame):  python tools/ncu_lines.py report.ncu-rep [kernel-substring] [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file = cur_fn = None
agg = {}
for r in csv.reader(txt.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        cur_fn = r[1]
    elif r[0] not in ("", "Line No", "Kernel Name") and sub in (cur_fn or ""):
        try:
            ln, s, ie = int(r[0]), int(r[6]), int(r[7])
        except ValueError:
            continue
        a = agg.setdefault((cur_file, ln), [0, 0, r[1][:100]])
        a[0] += s
        a[1] += ie
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
print(f"samples {tot} instructions {toti}")
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{k[0]:18s}{k[1]:5d} samp={a[0] / tot * 100:5.1f}% inst={a[1] / toti * 100:5.1f}%  {a[2]}")
