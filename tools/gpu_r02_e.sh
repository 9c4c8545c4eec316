#!/bin/bash
# round 2, call E (1 GPU): the final single-GPU evidence -- whole -m gpu suite, bench line with the
# pure-Python reference timed beside it, the reference arm, the config-2 size sweep, and the ncu
# launch lists / full captures behind profiles/r02_*
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 -p no:cacheprovider > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
tail -4 gpurun_out/e_pytest.log
timeout 600 python bench.py > gpurun_out/e_bench_n1.json 2> gpurun_out/e_bench_n1.err; echo "bench rc=$?" >> gpurun_out/e_bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/e_bench_ref.json 2> gpurun_out/e_bench_ref.err
timeout 600 python tools/gpu_ntt_sweep.py > gpurun_out/e_ntt_size_sweep.txt 2>&1
python tools/gpu_profile_proof.py > gpurun_out/e_proof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/e_launches_proof_2p20.csv python tools/gpu_profile_proof.py > gpurun_out/e_ncu_proof.log 2>&1
echo "ncu proof rc=$?"
python tools/gpu_profile_commit.py > gpurun_out/e_commit_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:merkle_leaf_pairs_cols_kernel -s 2 -c 1 -f -o gpurun_out/e_prof_merkle python tools/gpu_profile_commit.py > gpurun_out/e_ncu_merkle.log 2>&1
echo "ncu merkle rc=$?"
python tools/gpu_profile_commit.py > gpurun_out/e_commit_plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/e_launches_commit.csv python tools/gpu_profile_commit.py > gpurun_out/e_ncu_commit.log 2>&1
echo "ncu commit rc=$?"
ls -la gpurun_out | tail -15
