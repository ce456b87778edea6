#!/bin/bash
# same-box A/B of two library builds at a given batch size / kernel knob: tools/ab_n.sh libA.so libB.so N [PCG_WS] [rounds]
A=$1; B=$2; N=$3; W=${4:--1}; R=${5:-2}
for r in $(seq $R); do for L in $A $B; do
  if [ "$W" = "-1" ]; then unset PCG_WS; else export PCG_WS=$W; fi
  PCG_LIB=$PWD/$L python bench.py --steps 10 --warmup 3 --no-cpu-baseline --n-envs $N 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$L n=$N round $r kernel_ms=%.4f frac=%.3f %s'%(d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['kernel'][:24]))"
done; done
