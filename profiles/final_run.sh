set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/s2_pytest_gpu.txt
python bench.py > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s2_bench_ref.json 2>> gpurun_out/s2_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s2_launches.csv python profiles/profile_frame.py 2 > gpurun_out/s2_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:onesweep -s 24 -c 6 -o gpurun_out/s2_final_onesweep python profiles/profile_frame.py 3 > gpurun_out/s2_ncu.log 2>&1
cat gpurun_out/s2_pytest_gpu.txt; head -c 300 gpurun_out/s2_bench.json
