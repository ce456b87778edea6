#!/bin/bash
# round-2 GPU session E: full GPU suite on the shipped build, config 5, per-model numbers, ncu captures per kernel family
mkdir -p gpurun_out
T="timeout -k 5"
{
echo "== pytest -m gpu"
$T 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "== config 5 (1 GPU)"
$T 300 python bench.py --config 5 --steps 10 --warmup 3 2>&1 | tail -3
echo "== per-model"
$T 600 python tools/bench_models.py --steps 5 2>&1 | tail -14
echo "== headline bench line"
$T 400 python bench.py --steps 20 --warmup 5 2>&1 | tail -1
} > gpurun_out/r2_e.log 2>&1
# ncu: one full capture per kernel family (one launch each)
cap() {  # name, kernel regex, bench_models index
  $T 240 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o gpurun_out/r2_$1 -f \
     python tools/bench_models.py --only $3 --steps 2 --warmup 2 > gpurun_out/r2_e_ncu_$1.log 2>&1
}
cap fast_first_order_65k fast_rollout 0
cap fast_first_order_1m fast_rollout 1
cap generic_adaptive_cstr rollout_kernel 6
cap fast_four_tank rollout 7
cap extraction_adaptive rollout_kernel 10
tail -40 gpurun_out/r2_e.log
