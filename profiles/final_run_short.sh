# profiles/final_run.sh without the GPU test suite, most valuable outputs first (for the end of a round's GPU budget)
tag=${1:-r2}
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
ncu --set full --clock-control none --import-source on -s 42 -c 14 -o gpurun_out/${tag}_frame python profiles/profile_frame.py 1 > gpurun_out/${tag}_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-config5 > gpurun_out/${tag}_launch.log 2>&1
python profiles/configs_bench.py > gpurun_out/${tag}_configs.log 2>&1
cp gpurun_out/configs_r1.json gpurun_out/${tag}_configs.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
head -c 300 gpurun_out/${tag}_bench.json; tail -2 gpurun_out/${tag}_ncu.log
