#!/bin/bash
# round 2, call C (N GPUs, N = $1): multi-GPU parity under pytest (torchrun workers), then the
# bench line at N the way the driver launches it, and the reference arm under torchrun
N=${1:-2}
KEXPR=${2:-"$N"}   # which world sizes of tests/test_gpu_multi.py to run, e.g. "8" or "2 or 4"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/c${N}_gpus.txt 2>&1
nvidia-smi topo -m >> gpurun_out/c${N}_gpus.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -k "$KEXPR" --durations=5 -p no:cacheprovider > gpurun_out/c${N}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c${N}_pytest.log
tail -12 gpurun_out/c${N}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/c${N}_bench.json 2> gpurun_out/c${N}_bench.err; echo "bench rc=$?" >> gpurun_out/c${N}_bench.err
tail -c 1500 gpurun_out/c${N}_bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/c${N}_bench_ref.json 2> gpurun_out/c${N}_bench_ref.err
head -c 600 gpurun_out/c${N}_bench_ref.json
