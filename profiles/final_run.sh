# Round-end capture on one B200: GPU tests, bench (own + reference arm), all configs with full-size parity and stage times,
# ncu launch list of the bench command, ncu --set full of every kernel of one frame. Outputs under gpurun_out/<tag>_*.
# usage (under gpurun): bash profiles/final_run.sh r2
tag=${1:-r2}
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${tag}_pytest_gpu.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
python profiles/configs_bench.py > gpurun_out/${tag}_configs.log 2>&1
cp gpurun_out/configs_r1.json gpurun_out/${tag}_configs.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-config5 > gpurun_out/${tag}_launch.log 2>&1
ncu --set full --clock-control none --import-source on -s 42 -c 14 -o gpurun_out/${tag}_frame python profiles/profile_frame.py 1 > gpurun_out/${tag}_ncu.log 2>&1
cat gpurun_out/${tag}_pytest_gpu.txt; head -c 400 gpurun_out/${tag}_bench.json; tail -2 gpurun_out/${tag}_ncu.log
python profiles/sass_evidence.py > gpurun_out/${tag}_sass_opcodes.txt 2>/dev/null
