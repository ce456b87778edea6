#!/bin/bash
# round-2 GPU session W (8 GPUs): headline (quick line: device-resident only) and config 5 with the third session's kernels
mkdir -p gpurun_out
T="timeout -k 5"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
{
echo "== headline, 8 GPUs (quick)"
$T 120 $TR bench.py --gpus 8 --steps 20 --warmup 5 --quick 2>&1 | tail -1
echo "== config 5, 8 GPUs"
$T 200 $TR bench.py --config 5 --gpus 8 --steps 10 --warmup 3 2>&1 | tail -1
} > gpurun_out/r2_w8.log 2>&1
tail -6 gpurun_out/r2_w8.log | cut -c1-3500
