#!/usr/bin/env python
"""Per-CUDA-source-line stall samples of an .ncu-rep captured with --import-source on: python tools/ncu_lines.py REP [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
agg = []; cur = None
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if cur and len(r) > 7 and r[0].strip().isdigit() and r[2] == "-":
        try: agg.append((int(r[4]), int(r[7]), cur, int(r[0]), r[1].strip()[:100]))
        except ValueError: pass
tot = sum(a[0] for a in agg); agg.sort(reverse=True)
print("total samples", tot)
for a in agg[:top]:
    print(f"{a[0]:7d} {100 * a[0] / tot:5.1f}% inst={a[1]:10d} {a[2]}:{a[3]}  {a[4]}")
