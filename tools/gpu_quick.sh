#!/bin/bash
# quick GPU check after a kernel change: NTT/LDE parity tests, then headline timings
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ntt.py tests/test_gpu_edges.py tests/test_gpu_merkle_fri.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/quick.json 2> gpurun_out/quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/quick.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'extra', {k: d['extra'][k] for k in ('lde_merkle_commit_ms_64x2^18_x8', 'lde_ms', 'merkle_ms', 'stark_proof_s_fib_2^20_steps_x8')})
PY
