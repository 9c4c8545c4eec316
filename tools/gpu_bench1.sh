#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 6 -c 2 -o gpurun_out/prof_ntt python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_full.log
