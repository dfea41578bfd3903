"""Hottest SASS instructions (stall samples) of one launch of an .ncu-rep captured with --import-source on, with a few
instructions of context: python profiles/ncu_hot.py <report> [launch] [min share %] [context]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
want = 2 * (int(sys.argv[2]) if len(sys.argv) > 2 else 0)   # the sass view prints two identical blocks per launch
share = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
ctx = int(sys.argv[4]) if len(sys.argv) > 4 else 4
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
k, hdr, out = -1, None, []
for r in csv.reader(io.StringIO(raw)):
    if r and r[0] == "Kernel Name":
        k += 1
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if k != want or hdr is None or len(r) != len(hdr):
        continue
    out.append((int(r[hdr.index("# Samples")]), r[1].strip(), int(r[hdr.index("Instructions Executed")])))
tot = sum(o[0] for o in out)
print("samples", tot, "instructions", sum(o[2] for o in out))
for i, o in enumerate(out):
    if o[0] >= tot * share / 100:
        print(f"---- #{i}: {o[0] / tot * 100:.1f} % of samples")
        for j in range(max(i - ctx, 0), min(i + 2, len(out))):
            print(f"   {j:5d} {out[j][0]:6d} {out[j][2]:9d}  {out[j][1][:110]}")
