"""One line per variant of a run_variants.sh output file: python profiles/vsum.py gpurun_out/<file>.jsonl"""
import json
import sys
for l in open(sys.argv[1]):
    if not l.startswith("{"):
        continue
    d = json.loads(l); s = d["stages_ms"]
    per_pass = (s["tile_sort"] - s["tile_sort_hist_plan"]) / max(s["tile_passes_run"], 1)
    print(f"{d['label']:10s} ok={d['sorted_ok'] and d['ranges_ok']} pipelined {d['pipelined_ms']:.4f} frame {s['frame']:.4f} pre {s['preprocess']:.4f} depth {s['depth_sort']:.4f} dup {s['duplicate']:.4f} "
          f"tile {s['tile_sort']:.4f} (pass {per_pass * 1e3:.1f} us = {d['pairs'] * 16 / per_pass / 1e6 / 6540.8 * 100:.1f} %) ranges {s['ranges']:.4f} blend {s['blend']:.4f}")
