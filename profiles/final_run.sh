# Round-end capture on one B200: GPU tests, bench (own + reference arm), all configs with full-size parity, ncu launch list
# of one frame, ncu --set full of every kernel of one frame. Outputs under gpurun_out/.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/s3_pytest_gpu.txt
python bench.py > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s3_bench_ref.json 2>> gpurun_out/s3_bench.err
python profiles/configs_bench.py > gpurun_out/s3_configs.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s3_launches.csv python profiles/profile_frame.py 2 > gpurun_out/s3_launch.log 2>&1
ncu --set full --clock-control none --import-source on -s 64 -c 16 -o gpurun_out/s3_frame python profiles/profile_frame.py 2 > gpurun_out/s3_ncu.log 2>&1
cat gpurun_out/s3_pytest_gpu.txt; head -c 300 gpurun_out/s3_bench.json; tail -2 gpurun_out/s3_ncu.log
