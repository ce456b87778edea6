#!/bin/bash
# round-2 GPU session F: full GPU suite, FP64 issue microbenchmark, config 5, per-model numbers, adaptive per-SM kernel capture
mkdir -p gpurun_out
T="timeout -k 5"
{
echo "== fp64 issue microbench"
$T 60 tools/fp64_issue_microbench
echo "== pytest -m gpu"
$T 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40
echo "== config 5 (1 GPU)"
$T 300 python bench.py --config 5 --steps 10 --warmup 3 2>&1 | tail -3
echo "== per-model"
$T 600 python tools/bench_models.py --steps 5 2>&1 | tail -14
} > gpurun_out/r2_f.log 2>&1
$T 240 ncu --set full --clock-control none --import-source on -k regex:wsm_rollout -s 2 -c 1 -o gpurun_out/r2_wsm_adaptive_cstr -f \
     python tools/bench_models.py --only 6 --steps 2 --warmup 2 > gpurun_out/r2_f_ncu_wsm_adaptive.log 2>&1
tail -60 gpurun_out/r2_f.log
