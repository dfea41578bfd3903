"""Shared-memory wavefronts per SASS opcode class of an .ncu-rep captured with --import-source on (read here, no GPU):
which instructions pay for bank conflicts. python profiles/ncu_smem.py <report> [kernel index among the report's launches]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
want = 2 * (int(sys.argv[2]) if len(sys.argv) > 2 else 0)  # two identical blocks per launch
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
k, hdr = -1, None
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for r in rows:
    if r and r[0] == "Kernel Name":
        k += 1   # every launch of the report has one block per view: with --print-source sass that is one block per launch
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if k != want or hdr is None or len(r) != len(hdr) or "L1 Wavefronts Shared" not in hdr:
        continue
    op = r[1].split()
    op = op[1] if op[0].startswith("@") else op[0]
    w, ideal = int(r[hdr.index("L1 Wavefronts Shared")] or 0), int(r[hdr.index("L1 Wavefronts Shared Ideal")] or 0)
    if w == 0:
        continue
    a = agg[op.rstrip(";")]
    a[0] += int(r[hdr.index("Instructions Executed")]); a[1] += w; a[2] += ideal; a[3] += int(r[hdr.index("# Samples")])
tot = sum(v[1] for v in agg.values())
print(f"{'opcode':28s} {'warp-inst':>10s} {'wavefronts':>11s} {'ideal':>10s} {'per inst':>8s} {'share':>6s} samples")
for op, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{op:28s} {v[0]:10d} {v[1]:11d} {v[2]:10d} {v[1] / max(v[0], 1):8.2f} {v[1] / tot * 100:5.1f}% {v[3]}")
print("total wavefronts", tot)
