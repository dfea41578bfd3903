python -m pytest tests/test_shared_frames_gpu.py -x -q 2>&1 | tail -5
for m in direct gather; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --config5-collect $m > gpurun_out/n2_$m.json 2> gpurun_out/n2_$m.err
python - <<P
import json
l=json.load(open("gpurun_out/n2_$m.json")); c=l.get("config5") or l["config"].get("config5") or {}
print("$m", l["value"], l["ms_per_step"], json.dumps({k:c.get(k) for k in ("views_per_s","ms_per_batch","collect","views_with_pixels","frames_repeated")}))
P
tail -2 gpurun_out/n2_$m.err
done
