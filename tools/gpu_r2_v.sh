#!/bin/bash
# round-2 GPU session V: final ncu captures -- shipped headline kernel (with the full_a gate), general per-SM variant
# (config 5's cstr part), four_tank and crystallization generic kernels after the branch-free arithmetic.
# Summaries are condensed on the box (the reports together exceed what gpurun copies back).
mkdir -p gpurun_out
T="timeout -k 5"
N="ncu --set full --clock-control none --import-source on"
cap() {  # name regex skip command...
  local name=$1 re=$2 skip=$3; shift 3
  $T 240 $N -k regex:$re -s $skip -c 1 -o /tmp/$name -f "$@" > gpurun_out/$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py /dev/stdin > gpurun_out/${name}_ncu_summary.txt 2>&1
  rm -f /tmp/$name.ncu-rep
}
cap r2_cstr_wsm_final wsm_rollout 4 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
cap r2_cstr_wsm_general_cfg5 wsm_rollout 2 python bench.py --config 5 --steps 2 --warmup 3
cap r2_four_tank_fastmath rollout_kernel 2 python tools/bench_models.py --only 7 --steps 2 --warmup 2
cap r2_crystallization_fastmath rollout_kernel 2 python tools/bench_models.py --only 8 --steps 2 --warmup 2
head -8 gpurun_out/r2_cstr_wsm_final_ncu_summary.txt gpurun_out/r2_cstr_wsm_general_cfg5_ncu_summary.txt
