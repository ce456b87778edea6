#!/bin/bash
# round-2 GPU session U: lane-asynchronous adaptive integrator warps in the per-SM kernel (make_env's default scheme)
mkdir -p gpurun_out
T="timeout -k 5"
{
echo "== pytest"
$T 500 python -u -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short --timeout 180 -x -p no:cacheprovider -k "adaptive or fixed_step_probe or tile_count" 2>&1 | tail -12
echo "== cstr, default adaptive scheme"
$T 200 python tools/bench_models.py --steps 5 --only 6 2>&1 | tail -1 | cut -c1-420
} > gpurun_out/r2_u.log 2>&1
tail -30 gpurun_out/r2_u.log | cut -c1-500
