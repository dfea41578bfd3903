python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
python - <<P
import json
l=json.load(open("gpurun_out/r2h_bench.json")); print(l["value"], l["timed_region_s"], l["e2e"]["value"], l["e2e"]["steps"], l["e2e"]["serial_latency_ms"], l["roofline"]["frac"], l["config5"]["views_per_s"])
P
bash profiles/sanitize.sh > gpurun_out/r2h_sanitizer.txt 2>&1; cat gpurun_out/r2h_sanitizer.txt
