#!/bin/bash
# A/B of an env knob on the headline NTT / LDE timings: gpu_ab.sh VAR v1 v2 ...
mkdir -p gpurun_out
var=$1; shift
for v in "$@"; do
  export $var=$v
  python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1])
print('$var=$v', 'Melem/s %.0f' % d['value'], 'ms %.3f' % d['ms_per_step'], {k: round(d['extra'][k], 3) for k in ('lde_merkle_commit_ms_64x2^18_x8', 'lde_ms', 'merkle_ms', 'stark_proof_s_fib_2^20_steps_x8')})
PY
done
