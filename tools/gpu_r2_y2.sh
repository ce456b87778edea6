#!/bin/bash
# round-2 GPU session Y (2 GPUs): the driver's launch line on the final build (full headline line incl. e2e and the
# fixed-step probe in every rank), config 5, and the smoke entry point
mkdir -p gpurun_out
T="timeout -k 5"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519"
{
echo "== smoke"
$T 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== headline, 2 GPUs (full line)"
$T 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-1200
echo "== reference arm under torchrun"
$T 300 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
echo "== config 5, 2 GPUs"
$T 200 $TR bench.py --config 5 --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-400
} > gpurun_out/r2_y2.log 2>&1
tail -12 gpurun_out/r2_y2.log
