#!/bin/bash
# round 2, call B (1 GPU): the tests touched since call A, then the ncu evidence for the bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_upstream.py tests/test_gpu_dist_emulated.py tests/test_gpu_air.py tests/test_gpu_merkle_fri.py -m gpu -q --durations=15 -p no:cacheprovider > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -15 gpurun_out/b_pytest.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras"
$CMD > gpurun_out/b_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/b_launches_bench.csv $CMD > gpurun_out/b_ncu_list.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/b_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel -s 6 -c 2 -f -o gpurun_out/b_prof_ntt $CMD > gpurun_out/b_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -12
