#!/bin/bash
# round 2, final single-GPU run: smoke, the whole -m gpu suite, the bench line (pure-Python reference
# timed beside it) and the reference arm with the driver's step counts
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/f_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --durations=8 -p no:cacheprovider > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -4 gpurun_out/f_pytest.log; tail -2 gpurun_out/f_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err; echo "bench rc=$?" >> gpurun_out/f_bench_n1.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
tail -2 gpurun_out/f_bench_n1.err
