#!/usr/bin/env python
"""Per-opcode and per-source-line executed-instruction totals of one kernel from an ncu report's source page.
  python tools/ncu_source_hot.py <rep> <launch-skip> [top]"""
import csv, subprocess, sys, collections
rep, skip = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
ops = collections.Counter(); lines = collections.Counter(); stall = collections.Counter()
cur = None; curline = None; fn = ""; seen = set()
for row in csv.reader(out.splitlines()):
    if not row: continue
    if row[0] == "File Path": cur = row[1].split("/")[-1]; continue
    if row[0] == "Function Name": fn = row[1]; continue
    if row[0] == "Line No": continue
    if len(row) < 8: continue
    if row[0] != "":          # source row
        curline = (cur, row[0], row[1].strip()[:90]); continue
    if row[2].startswith("0x"):
        try: n = int(row[7]); s = int(row[6])
        except ValueError: continue
        if row[2] in seen: continue
        seen.add(row[2])
        op = row[3].split()[0]
        if op.startswith("@"): op = row[3].split()[1]
        ops[op] += n; lines[curline] += n; stall[curline] += s
tot = sum(ops.values())
print(fn[:100]); print("warp instructions executed:", tot)
for op, n in ops.most_common(top): print(f"  {op:28s} {n:10d} {100*n/tot:5.1f}%")
print("-- by source line (instructions, %, stall samples)")
for k, n in lines.most_common(top): print(f"  {k[0]}:{k[1]:>4s} {n:10d} {100*n/tot:5.1f}% {stall[k]:6d}  {k[2]}")
