#!/bin/bash
# round-2 final check: the driver's round-end sequence on the final build
mkdir -p gpurun_out
T="timeout -k 5"
{
echo "== pytest -m gpu (all)"
$T 1200 python -u -m pytest tests -x -q -m gpu --tb=short -p no:cacheprovider -W ignore::RuntimeWarning 2>&1 | tail -6
echo "== smoke"
$T 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench (default)"
$T 400 python bench.py 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value %.4g ms_per_step %.4f frac %.3f kernel_ms %.4f e2e %.4g launches %s clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'], d.get('gpu_launches'), d.get('clocks', {}).get('sm_mhz')))
print('cpu_baseline', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['kind'])
"
} > gpurun_out/r2_final.log 2>&1
tail -14 gpurun_out/r2_final.log | cut -c1-400
