#!/bin/bash
# round-2 GPU session B: full GPU suite on the ICDF-noise / table-exp build + same-box A/B of the three libraries
mkdir -p gpurun_out
{
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "== A/B kernel time (65 536 envs): round-1 library, table exp, table exp + ICDF noise + staged setpoints"
for r in 1 2 3; do for L in ab/lib_r1.so ab/lib_T.so ab/lib_I.so; do
  PCG_LIB=$PWD/$L python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$L round $r kernel_ms=%.4f frac=%.3f e2e=%.3e'%(d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value']))"
done; done
echo "== sweep (new lib)"
SWEEP_N="18944 37888 56832 65536 75776 131072 262144" SWEEP_WS="auto" SWEEP_NOISE="philox off" PCG_LIB=$PWD/ab/lib_I.so bash tools/sweep_envs.sh
} > gpurun_out/r2_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wsm_rollout -s 4 -c 1 -o gpurun_out/r2_wsm_I -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b_ncu.log 2>&1
tail -60 gpurun_out/r2_b.log
