"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel for the LAST `n` launches (one frame).
python profiles/launch_summary.py <launches.csv> <launches per frame>"""
import collections
import csv
import sys

path, per_frame = sys.argv[1], int(sys.argv[2])
rows = [r for r in csv.reader(open(path)) if len(r) > 5 and r[0].isdigit()]
hdr = next(r for r in csv.reader(open(path)) if "Kernel Name" in r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
frame = rows[-per_frame:]
agg = collections.OrderedDict()
for r in frame:
    name = r[ki].split("(")[0].replace("void ", "").replace("tpdcu::", "")
    if "<" in r[ki].split("(")[0]:
        name = r[ki].split("(")[0].replace("void ", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) / 1e3
total = sum(a[1] for a in agg.values())
for name, (cnt, us) in agg.items():
    print(f"{name:44s} launches {cnt}  total {us:8.1f} us  mean {us / cnt:8.1f} us  share {us / total * 100:5.1f} %")
print(f"{'frame (sum of launches)':44s} launches {sum(a[0] for a in agg.values())}  total {total:8.1f} us")
