#!/bin/bash
# round-2 GPU session N: general variant of the per-SM kernel (disturbance rows, substeps): parity + config 5
mkdir -p gpurun_out
T="timeout -k 5"
{
echo "== pytest"
$T 500 python -u -m pytest tests/test_gpu_parity.py tests/test_gpu_mixed.py -q -m gpu --tb=short --timeout 120 -x -p no:cacheprovider -k "general_per_sm or mixed or config5 or per_env_setpoints or step_cases" 2>&1 | tail -25
echo "== config 5 (1 GPU)"
$T 300 python bench.py --config 5 --steps 10 --warmup 3 2>&1 | tail -1
} > gpurun_out/r2_n.log 2>&1
tail -40 gpurun_out/r2_n.log | cut -c1-3000
