#!/bin/bash
# round 2, call A (1 GPU): smoke, the whole -m gpu suite, the bench line and the reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
nproc >> gpurun_out/a_gpu.txt; free -g | head -2 >> gpurun_out/a_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/a_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --durations=30 -p no:cacheprovider > gpurun_out/a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/a_bench_n1.json 2> gpurun_out/a_bench_n1.err; echo "bench rc=$?" >> gpurun_out/a_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err
tail -5 gpurun_out/a_pytest_gpu.log; tail -3 gpurun_out/a_smoke.log; tail -c 600 gpurun_out/a_bench_n1.err
