#!/bin/bash
# round-2 GPU session J: direct-index exponential table in the per-SM kernel (CstrW): parity tests + same-box A/B
mkdir -p gpurun_out
T="timeout -k 5"
{
echo "== pytest (per-SM kernel parity)"
$T 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -k "straight_line or headline or adaptive or kernel or noise" 2>&1 | tail -15
echo "== A/B"
for r in 1 2; do for L in ab/lib_BASE.so pcgym_b200/libpcgym_b200.so ab/lib_F128.so ab/lib_W16.so; do
  PCG_LIB=$PWD/$L $T 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$L round $r kernel_ms=%.4f frac=%.3f value=%.4g %s'%(d['roofline']['kernel_ms'], d['roofline']['frac'], d['value'], d['roofline']['kernel'][:30]))"
done; done
echo "== adaptive default"
for L in ab/lib_BASE.so pcgym_b200/libpcgym_b200.so; do
  PCG_LIB=$PWD/$L $T 300 python tools/bench_models.py --steps 5 --only 6 2>&1 | tail -1 | cut -c1-400
done
} > gpurun_out/r2_j.log 2>&1
tail -40 gpurun_out/r2_j.log | cut -c1-300
