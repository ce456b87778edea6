#!/bin/bash
# round-2 GPU session G: full GPU suite (short tracebacks), fast-kernel cp.async staging A/B on the first-order rows
mkdir -p gpurun_out
T="timeout -k 5"
{
echo "== pytest -m gpu"
$T 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -90
echo "== per-model (first_order rows, four_tank, crystallization)"
$T 600 python tools/bench_models.py --steps 5 --only 0,1,2,3,7,8 2>&1 | tail -8
echo "== config 5 (1 GPU)"
$T 300 python bench.py --config 5 --steps 10 --warmup 3 2>&1 | tail -2
} > gpurun_out/r2_g.log 2>&1
tail -100 gpurun_out/r2_g.log
