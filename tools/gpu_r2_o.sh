#!/bin/bash
# round-2 GPU session O: branch-free sqrt / log / reciprocal in four_tank and crystallization: full GPU suite, config 5, per-model
mkdir -p gpurun_out
T="timeout -k 5"
{
echo "== pytest -m gpu (all)"
$T 900 python -u -m pytest tests -q -m gpu --tb=short --timeout 180 -p no:cacheprovider 2>&1 | tail -25
echo "== config 5 (1 GPU)"
$T 300 python bench.py --config 5 --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value %.4g  ms_per_step %.4f frac %.4f'%(d['value'], d['ms_per_step'], d['roofline']['frac']))
for k,v in d['parts_alone'].items(): print(' ', k, round(v['ms'],4), v['kernel'])
print('e2e', d['e2e'])
"
echo "== per-model"
$T 300 python tools/bench_models.py --steps 5 --only 7,8 2>&1 | tail -2 | cut -c1-420
} > gpurun_out/r2_o.log 2>&1
tail -50 gpurun_out/r2_o.log | cut -c1-600
