"""Summarise the per-tile trace of the onesweep pass kernel (profiles/micro/ws_trace.cu built from sort.cu with
-DTPDCU_WS_TRACE): python profiles/os_trace_summary.py gpurun_out/os_trace.txt"""
import sys
import numpy as np
lines = open(sys.argv[1]).read().splitlines()
print(lines[0])
a = np.array([[float(x) for x in l.split() if x != "|"] for l in lines if l and l[0].isdigit()])
t = a[:, 0]; start, agg, lbs, lbd, end, depth = a[:, 2], a[:, 4], a[:, 5], a[:, 6], a[:, 9], a[:, 10]
def st(n, x):
    print(f"{n:30s} mean {x.mean():7.0f} p10 {np.percentile(x, 10):7.0f} p50 {np.percentile(x, 50):7.0f} p90 {np.percentile(x, 90):7.0f} max {x.max():7.0f}")
print("tiles", len(a), "kernel ns", end.max())
st("start->aggregate (count)", agg - start); st("aggregate->look-back (rank)", lbs - agg); st("look-back", lbd - lbs); st("write-out", end - lbd)
st("tile life", end - start); st("look-back depth (rows)", depth)
ev = np.arange(0, end.max(), 4000)
print("tiles in flight every 4 us:", [int(((start <= x) & (end > x)).sum()) for x in ev])
