#!/bin/bash
# round-2 GPU session C: 1-env-per-lane integrators + cp.async action ring; A/B against the earlier builds
mkdir -p gpurun_out
{
echo "== pytest -m gpu"
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "== A/B kernel time (65 536 envs)"
for r in 1 2; do for L in ab/lib_T.so ab/lib_I.so ab/lib_E2.so ab/lib_E1.so; do
  PCG_LIB=$PWD/$L python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$L round $r kernel_ms=%.4f frac=%.3f graph_ms=%.4f e2e=%.3e reduced=%.3e enq_us=%.1f'%(d['roofline']['kernel_ms'], d['roofline']['frac'], d['ms_per_step'], d['e2e']['value'], d['e2e']['reduced']['value'], d['per_rank']['rows'][0][2]))"
done; done
echo "== sweep (E1)"
SWEEP_N="18944 37888 56832 65536 75776 131072" SWEEP_WS="auto" SWEEP_NOISE="philox off" PCG_LIB=$PWD/ab/lib_E1.so bash tools/sweep_envs.sh
echo "== full bench line"
python bench.py --steps 20 --warmup 5
} > gpurun_out/r2_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wsm_rollout -s 4 -c 1 -o gpurun_out/r2_wsm_E1 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2_c_ncu.log 2>&1
tail -40 gpurun_out/r2_c.log
