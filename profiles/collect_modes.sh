# usage (under gpurun --gpus N): bash profiles/collect_modes.sh N "direct push gather" — the 64-view batch with every way of collecting the frames
n=${1:-4}; modes=${2:-"direct push gather"}
[ "$n" -le 4 ] && python -m pytest tests/test_shared_frames_gpu.py -x -q 2>&1 | tail -3
for m in $modes; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline --config5-collect $m > gpurun_out/n${n}_$m.json 2> gpurun_out/n${n}_$m.err
python - <<P
import json
l=json.load(open("gpurun_out/n${n}_$m.json")); c=l.get("config5") or l["config"].get("config5") or {}
print("$m", l["value"], json.dumps({k:c.get(k) for k in ("views_per_s","ms_per_batch","collect","views_with_pixels","frames_repeated","own_frames_ms_per_rank_last_batch")}))
P
tail -1 gpurun_out/n${n}_$m.err
done
