#!/bin/bash
# A/B of the plain pass kernel compiled for 4 (default), 3 and 2 resident CTAs per SM (128 / 168 / 210
# registers per thread): same box, same run; headline NTT + the size sweep's key points
mkdir -p gpurun_out
for lib in "" _m3 _m2; do
  if [ -n "$lib" ]; then export STARKS_B200_LIB=$PWD/starks_b200/libstarks_b200$lib.so; else unset STARKS_B200_LIB; fi
  echo "== lib${lib:-default}" >> gpurun_out/minb_ab.txt
  python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('ntt 64x2^20 ms', d['ms_per_step'], 'Melem/s', d['value'], 'parity', d['parity'])" >> gpurun_out/minb_ab.txt
done
cat gpurun_out/minb_ab.txt
