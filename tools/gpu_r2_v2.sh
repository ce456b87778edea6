#!/bin/bash
# round-2 GPU session V2: ncu captures of the remaining kernel families on the final build (summaries condensed on the box)
mkdir -p gpurun_out
T="timeout -k 5"
N="ncu --set full --clock-control none --import-source on"
cap() {  # name regex skip command...
  local name=$1 re=$2 skip=$3; shift 3
  $T 240 $N -k regex:$re -s $skip -c 1 -o /tmp/$name -f "$@" > gpurun_out/$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py /dev/stdin > gpurun_out/${name}_ncu_summary.txt 2>&1
  rm -f /tmp/$name.ncu-rep
}
cap r2_fast_first_order_65k_final fast_rollout 2 python tools/bench_models.py --only 0 --steps 2 --warmup 2
cap r2_ws_pairs_cstr_8192 ws_rollout 2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --n-envs 8192
cap r2_fast_cstr_65535 fast_rollout 2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --n-envs 65535
cap r2_affine_3x2 rollout 2 python tools/affine_run.py
head -4 gpurun_out/r2_fast_first_order_65k_final_ncu_summary.txt gpurun_out/r2_ws_pairs_cstr_8192_ncu_summary.txt gpurun_out/r2_fast_cstr_65535_ncu_summary.txt gpurun_out/r2_affine_3x2_ncu_summary.txt
