#!/bin/bash
# round-2 GPU session M: per-SM kernel in several waves of CTAs for batches beyond 8 tiles per SM
mkdir -p gpurun_out
T="timeout -k 5"
{
echo "== pytest"
$T 400 python -u -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short --timeout 120 -x -p no:cacheprovider -k "split or three_straight or tile_count" 2>&1 | tail -8
echo "== sweep (auto = per-SM kernel in waves, ws=1 pair-per-CTA, ws=0 monolithic)"
SWEEP_N="75776 98304 131072 262144 524288" SWEEP_WS="auto 1 0" SWEEP_NOISE="philox" $T 500 bash tools/sweep_envs.sh 2>&1
} > gpurun_out/r2_m.log 2>&1
tail -30 gpurun_out/r2_m.log | cut -c1-200
