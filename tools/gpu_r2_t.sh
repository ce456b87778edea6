#!/bin/bash
# round-2 GPU session T: the round-end sequence -- full GPU suite, smoke, both bench arms, config 2 / 4 / 5, launch list
mkdir -p gpurun_out
T="timeout -k 5"
{
echo "== pytest -m gpu (all)"
$T 1200 python -u -m pytest tests -q -m gpu --tb=short --timeout 240 -p no:cacheprovider 2>&1 | tail -15
echo "== smoke"
$T 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (default)"
$T 400 python bench.py 2>&1 | tail -1
echo "== bench --impl reference"
$T 400 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-600
for c in 2 4 5; do echo "== bench --config $c"; $T 300 python bench.py --config $c --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-1500; done
} > gpurun_out/r2_t.log 2>&1
$T 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_t_ncu.log 2>&1
tail -40 gpurun_out/r2_t.log | cut -c1-2500
