#!/bin/bash
# round-2 GPU session R: lane-split kernel with the feed-lane trick / folded step; occupancy A/B (4, 5, 6 CTAs per SM)
mkdir -p gpurun_out
T="timeout -k 5"
{
echo "== pytest"
$T 400 python -u -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short --timeout 180 -x -p no:cacheprovider -k "lane_split or (batch_rollout and extraction) or (golden and extraction)" 2>&1 | tail -12
echo "== config 4, 1e-8: main (4 CTAs/SM), L5, L6"
for L in pcgym_b200/libpcgym_b200.so ab/lib_L5.so ab/lib_L6.so; do
  PCG_LIB=$PWD/$L $T 200 python tools/bench_models.py --steps 3 --only 10,15 2>&1 | tail -2 | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); print('$L', d['model'], d['ms_per_launch'], '%.4g' % d['env_steps_per_s'], d['rhs_evals_per_interval'], d['kernel'][:22])
"
done
} > gpurun_out/r2_r.log 2>&1
tail -30 gpurun_out/r2_r.log | cut -c1-300
