"""Executed-instruction histogram (by opcode, and by contiguous SASS block) from an ncu sass source-page csv."""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1]))); ki = int(sys.argv[2]) if len(sys.argv) > 2 else 0
kern = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = dict(name=r[1], hdr=None, rows=[]); kern.append(cur); continue
    if cur is None: continue
    if r and r[0] == "Address": cur["hdr"] = r; continue
    if cur["hdr"] and len(r) == len(cur["hdr"]): cur["rows"].append(r)
k = kern[ki]; h = {n: i for i, n in enumerate(k["hdr"])}
ie = h["Instructions Executed"]
tot = sum(int(r[ie] or 0) for r in k["rows"])
print("total warp-instr", tot, "static instr", len(k["rows"]))
c = Counter()
for r in k["rows"]:
    src = r[h["Source"]].split(); op = src[1] if src[0].startswith('@') else src[0]
    c[op.split('.')[0]] += int(r[ie] or 0)
print("  ".join(f"{op}={100*n/tot:.1f}%" for op, n in c.most_common(24)))
# blocks: runs of instructions with identical executed count
blocks = []; start = 0
R = k["rows"]
for i in range(1, len(R) + 1):
    if i == len(R) or R[i][ie] != R[start][ie]:
        blocks.append((start, i, int(R[start][ie] or 0))); start = i
blocks.sort(key=lambda b: -(b[1] - b[0]) * b[2])
for s, e, n in blocks[:int(sys.argv[3]) if len(sys.argv) > 3 else 12]:
    ops = Counter()
    for r in R[s:e]:
        src = r[h["Source"]].split(); op = src[1] if src[0].startswith('@') else src[0]; ops[op.split('.')[0]] += 1
    print(f"[{R[s][h['Address']][-5:]}..{R[e-1][h['Address']][-5:]}] {e-s:4d} instr x {n:8d} = {100*(e-s)*n/tot:5.1f}%  " +
          " ".join(f"{o}:{m}" for o, m in ops.most_common(8)))
