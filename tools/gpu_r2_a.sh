#!/bin/bash
# round-2 GPU session A: correctness of the table-exp integrator + same-box A/B against the round-1 library
mkdir -p gpurun_out
{
echo "== parity subset"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "three_straight or split_rollouts or headline or batch_rollout_vs_oracle and cstr or golden_traj" 2>&1 | tail -5
echo "== A/B kernel time (65 536 envs)"
bash tools/ab.sh ab/lib_r1.so ab/lib_T.so 3
echo "== noise off"
for L in ab/lib_r1.so ab/lib_T.so; do PCG_LIB=$PWD/$L python bench.py --steps 20 --warmup 5 --no-cpu-baseline --noise-mode off 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$L noise off kernel_ms=%.4f frac=%.3f'%(d['roofline']['kernel_ms'], d['roofline']['frac']))"; done
echo "== sweep (new lib, auto kernel choice + forced)"
SWEEP_N="18944 37888 56832 65536 75776 131072 262144" SWEEP_WS="auto 1" SWEEP_NOISE="philox" PCG_LIB=$PWD/ab/lib_T.so bash tools/sweep_envs.sh
} > gpurun_out/r2_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wsm_rollout -s 4 -c 1 -o gpurun_out/r2_wsm_T -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_a_ncu.log 2>&1
tail -40 gpurun_out/r2_a.log
