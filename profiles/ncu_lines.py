"""Per-source-line instruction and stall-sample shares of an .ncu-rep captured with --import-source on.
python profiles/ncu_lines.py <report> [top N] [launch index within the report; default: all launches together]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
want = int(sys.argv[3]) if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur, hdr = None, None
first_file, launch = None, -1
agg = collections.defaultdict(lambda: [0, 0, "", 0])
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        if first_file is None:
            first_file = r[1]
        if r[1] == first_file:
            launch += 1   # the per-file blocks of every launch start over with the same file
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if want is not None and launch != want:
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        try:
            n, s, t = int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")]), int(r[hdr.index("Thread Instructions Executed")])
        except ValueError:
            continue
        key = (cur, int(r[0]))
        agg[key][0] += n
        agg[key][1] += s
        agg[key][2] = r[1]
        agg[key][3] += t
tot = sum(v[0] for v in agg.values())
ts = sum(v[1] for v in agg.values())
print("total warp-instructions", tot, "samples", ts)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[0][:12]:12s}:{k[1]:4d} inst {v[0] / tot * 100:5.1f}% lanes {v[3] / max(v[0], 1):4.1f} samples {v[1] / max(ts, 1) * 100:5.1f}%  {v[2].strip()[:95]}")
