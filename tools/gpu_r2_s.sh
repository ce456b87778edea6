#!/bin/bash
# round-2 GPU session S: PI ("stabilised") step-size controller in the lane-split kernel: beta = 0 / 0.04 / 0.08
mkdir -p gpurun_out
T="timeout -k 5"
{
for L in pcgym_b200/libpcgym_b200.so ab/lib_B40.so ab/lib_B80.so; do
  echo "== $L"
  PCG_LIB=$PWD/$L $T 200 python tools/bench_models.py --steps 3 --only 10,15 2>&1 | tail -2 | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); print(d['model'], d['ms_per_launch'], '%.4g' % d['env_steps_per_s'], d['rhs_evals_per_interval'], d['kernel'][:22])
"
  PCG_LIB=$PWD/$L $T 200 python tools/extraction_tol.py 2>&1 | tail -5 | head -3
done
} > gpurun_out/r2_s.log 2>&1
tail -30 gpurun_out/r2_s.log | cut -c1-300
