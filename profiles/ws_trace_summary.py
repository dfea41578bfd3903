"""Summarise profiles/micro/ws_trace output: python profiles/ws_trace_summary.py gpurun_out/ws_trace.txt"""
import sys
import numpy as np
lines = open(sys.argv[1]).read().splitlines()
print(lines[0])
a = np.array([[float(x) for x in l.split() if x != "|"] for l in lines if l and l[0].isdigit()])
draw, keys, agg, lbs, lbd, rank, bases, wr = [a[:, i] for i in range(2, 10)]
depth, ret = a[:, 10], a[:, 11]
def st(name, x):
    print(f"{name:36s} mean {x.mean():8.0f}  p10 {np.percentile(x, 10):8.0f} p50 {np.percentile(x, 50):8.0f} p90 {np.percentile(x, 90):8.0f} max {x.max():8.0f}")
print("tiles", len(a), "last pass ns", wr.max())
st("draw->keys_in (TMA wait)", keys - draw); st("keys_in->agg (count+per-bin)", agg - keys); st("lb_start->lb_done (look-back)", lbd - lbs)
st("agg->rank_done (rank+scatter)", rank - agg); st("rank_done->bases_in (wait helper)", bases - rank); st("bases_in->write_done", wr - bases)
st("tile life draw->write_done", wr - draw); st("look-back depth (rows)", depth); st("row re-fetches", ret)
