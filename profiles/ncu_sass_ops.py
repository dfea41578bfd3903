"""Opcode mix, stall samples and shared-memory wavefronts of one kernel launch of an .ncu-rep (captured with
--import-source on), split at block barriers. python profiles/ncu_sass_ops.py <report> <launch index> [min inst share]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, which = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
kern, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []
        kern.append(cur)
        continue
    if cur is not None:
        cur.append(r)
kk = kern[which]
hdr, body = kk[0], kk[1:]
iS, iI, iW, iWi = (hdr.index(x) for x in ("# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal"))
tot, toti = sum(int(r[iS]) for r in body), sum(int(r[iI]) for r in body)
totw = sum(int(r[iW] or 0) for r in body)
print("launches in report", len(kern), "| samples", tot, "warp-inst", toti, "smem wavefronts", totw)
seg, start = 0, 0
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
s = i = w = 0
for idx, r in enumerate(body):
    s += int(r[iS]); i += int(r[iI]); w += int(r[iW] or 0)
    op = re.sub(r"^@!?U?P\d+\s+", "", r[1].strip()).split()[0]
    x = agg[op]
    x[0] += int(r[iI]); x[1] += int(r[iS]); x[2] += int(r[iW] or 0); x[3] += int(r[iWi] or 0)
    if "BAR.SYNC" in r[1] or idx == len(body) - 1:
        print(f"segment {seg} [{start}-{idx}] samples {100 * s / tot:5.1f}% inst {100 * i / toti:5.1f}% wavefronts {w}")
        for op, x in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            if x[2] or x[0] > toti * 0.01:
                print(f"    {op:26s} inst {x[0]:9d} samples {x[1]:5d} wavefronts {x[2]:8d} ideal {x[3]:8d}")
        seg += 1; start = idx + 1; s = i = w = 0
        agg = collections.defaultdict(lambda: [0, 0, 0, 0])
