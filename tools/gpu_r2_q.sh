#!/bin/bash
# round-2 GPU session Q: ncu capture of the lane-split extraction kernel (config 4) and of the general per-SM variant
mkdir -p gpurun_out
T="timeout -k 5"
$T 300 ncu --set full --clock-control none --import-source on -k regex:lanes_rollout -s 2 -c 1 -o gpurun_out/r2_lanes_extraction -f \
     python tools/bench_models.py --only 10 --steps 2 --warmup 2 > gpurun_out/r2_q_ncu_lanes.log 2>&1
tail -3 gpurun_out/r2_q_ncu_lanes.log | cut -c1-300
ls -la gpurun_out/r2_lanes_extraction.ncu-rep
