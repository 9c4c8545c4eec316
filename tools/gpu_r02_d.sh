#!/bin/bash
# round 2, call D (1 GPU): canary bounds tests + everything that hashes, then the bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=12 -p no:cacheprovider > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
tail -22 gpurun_out/d_pytest.log
timeout 600 python bench.py --no-pyref > gpurun_out/d_bench_n1.json 2> gpurun_out/d_bench_n1.err; echo "bench rc=$?" >> gpurun_out/d_bench_n1.err
tail -3 gpurun_out/d_bench_n1.err
