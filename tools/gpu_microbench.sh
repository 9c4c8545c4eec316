#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/clocks_mb.csv &
SMI=$!
python - <<'PY' 2>&1 | tee gpurun_out/microbench.log
import sys; sys.path.insert(0, '.')
from starks_b200 import Engine
e = Engine(0)
names = {0:"IMAD",1:"IMAD.WIDE+IADD3+IADD3.X (split)",2:"IADD3",3:"IMAD.HI(+MOV)",4:"IADD3+LOP3+SHF",5:"field_mul",6:"butterfly",
         7:"IMAD+IADD3 indep",8:"IMAD.WIDE+2.5 IADD3",9:"IADD3.X carry chain",10:"IMAD.WIDE.X rows",
         11:"DFMA (fp64 pipe)",12:"DFMA + IMAD.WIDE indep",13:"DFMA + 2 IADD3 indep"}
for w in range(14):
    iters = 20000 if w not in (5,6) else 2000
    best = None
    for rep in range(3):
        ms, ops = e.microbench(w, iters)
        r = ops / (ms * 1e-3)
        best = max(best or 0, r)
    print("%-34s %.3f Gop/s  (%.3f ms)" % (names[w], best / 1e9, ms))
PY
kill $SMI
sort gpurun_out/clocks_mb.csv | uniq -c | sort -rn | head -5
