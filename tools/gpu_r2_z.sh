#!/bin/bash
# round-2 GPU session Z: compute-sanitizer over every kernel family (memcheck, racecheck, synccheck)
mkdir -p gpurun_out
T="timeout -k 5"
{
echo "== plain run"
$T 120 python tools/sanitize_run.py 2>&1 | tail -9
for tool in memcheck synccheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  $T 420 compute-sanitizer --tool $tool python tools/sanitize_run.py 2>&1 | grep -v "^$" | tail -14
done
} > gpurun_out/r2_z.log 2>&1
tail -60 gpurun_out/r2_z.log | cut -c1-220
