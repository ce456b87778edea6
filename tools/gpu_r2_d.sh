#!/bin/bash
# round-2 GPU session D: warp layouts with setmaxnreg + back-off waits; A/B against session C's build (lib_E1)
mkdir -p gpurun_out
{
echo "== parity (main lib = variant B)"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "straight_line_kernel or headline or three_kernel or split_rollouts or shard" 2>&1 | tail -8
echo "== A/B kernel time (65 536 envs)"
for r in 1 2; do for v in E1 F A B C D E G; do
  PCG_LIB=$PWD/ab/lib_$v.so python bench.py --steps 20 --warmup 5 --quick 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v round $r kernel_ms=%.4f graph_ms=%.4f frac=%.3f status=%d'%(d['kernel_ms'], d['graph_ms_per_step'], d['frac'], d['status_max']))"
done; done
echo "== other batch sizes"
for n in 18944 37888 56832 75776; do for v in E1 B C; do
  PCG_LIB=$PWD/ab/lib_$v.so python bench.py --steps 20 --warmup 5 --quick --n-envs $n 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v n=$n kernel_ms=%.4f frac=%.3f %s'%(d['kernel_ms'], d['frac'], d['kernel'][:28]))"
done; done
} > gpurun_out/r2_d.log 2>&1
for v in B C; do
PCG_LIB=$PWD/ab/lib_$v.so ncu --set full --clock-control none --import-source on -k regex:wsm_rollout -s 4 -c 1 -o gpurun_out/r2_wsm_$v -f python bench.py --steps 3 --warmup 3 --quick --no-graph > gpurun_out/r2_d_ncu_$v.log 2>&1
done
tail -50 gpurun_out/r2_d.log
