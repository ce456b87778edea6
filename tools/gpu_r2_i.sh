#!/bin/bash
# round-2 GPU session I: full GPU suite with the Rosenbrock scheme, config 5 on one GPU, stiff-integrator timings
mkdir -p gpurun_out
T="timeout -k 5"
{
echo "== pytest -m gpu"
$T 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -60
echo "== config 5 (1 GPU)"
$T 300 python bench.py --config 5 --steps 10 --warmup 3 2>&1 | tail -2
echo "== per-model (extraction explicit / rodas4, cstr adaptive / rodas4)"
$T 900 python tools/bench_models.py --steps 3 --only 6,10,11,12,13,14 2>&1 | tail -8
} > gpurun_out/r2_i.log 2>&1
tail -80 gpurun_out/r2_i.log | cut -c1-600
