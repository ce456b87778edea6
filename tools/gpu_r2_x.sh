#!/bin/bash
# round-2 GPU session X: PCG_F_EPISODE_START (no memset between back-to-back episodes) + programmatic dependent launch
# of the per-SM kernel: full GPU suite, then the headline with PDL off / on
mkdir -p gpurun_out
T="timeout -k 5"
{
echo "== pytest -m gpu (all)"
$T 1200 python -u -m pytest tests -q -m gpu --tb=short --timeout 240 -p no:cacheprovider -W ignore::RuntimeWarning 2>&1 | tail -8
echo "== headline: PCG_PDL=0 / 1, two rounds"
for r in 1 2; do for p in 0 1; do
  PCG_PDL=$p $T 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('PDL=$p round $r ms_per_step=%.4f kernel_ms=%.4f frac=%.3f value=%.4g'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['value']))"
done; done
echo "== config 5"
$T 200 python bench.py --config 5 --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('config 5 value %.4g ms %.4f'%(d['value'], d['ms_per_step']))"
} > gpurun_out/r2_x.log 2>&1
tail -20 gpurun_out/r2_x.log | cut -c1-300
