#!/bin/bash
# round-2 GPU session P: lane-split, lane-asynchronous adaptive kernel for multistage_extraction
mkdir -p gpurun_out
T="timeout -k 5"
{
echo "== pytest"
$T 600 python -u -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short --timeout 180 -x -p no:cacheprovider -k "lane_split or extraction or golden or repeated_steps" 2>&1 | tail -25
echo "== config 4 rows (lanes = auto)"
$T 400 python tools/bench_models.py --steps 3 --only 10,14,15 2>&1 | tail -3 | cut -c1-460
echo "== tolerance vs accuracy"
$T 300 python tools/extraction_tol.py 2>&1 | tail -6
} > gpurun_out/r2_p.log 2>&1
tail -60 gpurun_out/r2_p.log | cut -c1-600
