#!/bin/bash
# Round-1 (second session) evidence run on one B200: tests, bench (both arms), launch lists,
# ncu --set full of the transform pass kernel, proof launch list.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_r01b.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01b.json 2> gpurun_out/bench_r01b.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01b_reference.json 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 6 -c 2 -o gpurun_out/prof_ntt_r01b python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_full.log 2>&1
echo "ntt full rc=$?"
python tools/gpu_profile_proof.py > gpurun_out/proof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_proof_r01b.csv python tools/gpu_profile_proof.py > gpurun_out/ncu_proof.log 2>&1
echo "proof list rc=$?"; tail -1 gpurun_out/proof_plain.log
cat gpurun_out/bench_r01b.json
