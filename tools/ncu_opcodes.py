"""Opcode histogram (executed warp instructions) and top stall locations of each kernel in an .ncu-rep captured with
--import-source on.  usage: ncu_opcodes.py report.ncu-rep [kernel-substring]"""
import csv
import io
import subprocess
import sys
from collections import Counter

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
want = sys.argv[2] if len(sys.argv) > 2 else ""
blocks, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = [line]
        blocks.append(cur)
    elif cur is not None:
        cur.append(line)
for b in blocks:
    name = next(csv.reader([b[0]]))[1]
    if want not in name:
        continue
    rows = list(csv.DictReader(io.StringIO("\n".join(b[1:]))))
    ops, stalls = Counter(), []
    tot = 0
    for r in rows:
        n = int(r["Instructions Executed"] or 0)
        op = r["Source"].split()[0] if r["Source"].split() else "?"
        if op.startswith("@"):
            op = r["Source"].split()[1]
        ops[op.split(".")[0]] += n
        tot += n
        stalls.append((int(r["Warp Stall Sampling (All Samples)"] or 0), r["Source"].strip()[:70], n))
    print("==", name[:90], "| static", len(rows), "| executed", tot)
    print("   " + "  ".join(f"{o}:{c / tot * 100:.1f}%" for o, c in ops.most_common(14)))
    samp = sum(s for s, _, _ in stalls) or 1
    for s, src, n in sorted(stalls, reverse=True)[:8]:
        print(f"   {s / samp * 100:5.1f}% samples  x{n:9d}  {src}")
