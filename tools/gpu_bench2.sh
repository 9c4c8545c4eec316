#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench rc=$?"; cat gpurun_out/bench_r01.json; tail -3 gpurun_out/bench_r01.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_reference.json 2>&1; cat gpurun_out/bench_r01_reference.json
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 6 -c 2 -o gpurun_out/prof_ntt_r01 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/ncu_full.log
