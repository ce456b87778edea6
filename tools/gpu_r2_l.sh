#!/bin/bash
# round-2 GPU session L: full_a gate -- the stalled size, the per-SM kernel parity tests, same-box A/B against the round's base
mkdir -p gpurun_out
T="timeout -k 5"
{
echo "== stalled size"
$T 200 python tools/dbg_hang.py "" 2>&1 | cut -c1-200
echo "== pytest (per-SM kernel parity)"
$T 600 python -u -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short --timeout 120 -x -p no:cacheprovider \
  -k "straight_line or headline or adaptive or kernel or noise or split" 2>&1 | tail -15
echo "== A/B"
for r in 1 2 3; do for L in ab/lib_BASE.so pcgym_b200/libpcgym_b200.so; do
  PCG_LIB=$PWD/$L $T 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$L round $r kernel_ms=%.4f frac=%.3f value=%.4g %s'%(d['roofline']['kernel_ms'], d['roofline']['frac'], d['value'], d['roofline']['kernel'][:30]))"
done; done
} > gpurun_out/r2_l.log 2>&1
tail -40 gpurun_out/r2_l.log | cut -c1-300
