#!/bin/bash
# first GPU session: parity tests, microbenchmarks
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
python - <<'PY' 2>&1 | tee gpurun_out/microbench.log
import sys; sys.path.insert(0, '.')
from starks_b200 import Engine
e = Engine(0)
names = ["IMAD","IMAD.WIDE","IADD3","IMAD.WIDE+IADD3","IADD3+LOP3+SHF","field_mul","butterfly"]
for w in range(7):
    iters = 20000 if w < 5 else 2000
    best = None
    for rep in range(3):
        ms, ops = e.microbench(w, iters)
        r = ops / (ms * 1e-3)
        best = max(best or 0, r)
    print("%-18s %.3f Gop/s  (%.3f ms)" % (names[w], best / 1e9, ms))
PY
