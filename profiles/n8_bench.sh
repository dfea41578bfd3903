# usage (under gpurun --gpus N): bash profiles/n8_bench.sh N tag
n=${1:-8}; tag=${2:-r2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
python - <<P
import json
l=json.load(open("gpurun_out/${tag}_bench_n$n.json")); c=l.get("config5") or l["config"].get("config5") or {}
print(l["n_gpus"], l["value"], l["ms_per_step"], l.get("e2e"), json.dumps({k:c.get(k) for k in ("views_per_s","ms_per_batch","collect","views_with_pixels","frames_repeated")}))
P
tail -3 gpurun_out/${tag}_bench_n$n.err
