#!/bin/bash
# A/B two builds of the library on the same box: tools/ab.sh libA.so libB.so [rounds]   (profiling aid)
A=$1; B=$2; R=${3:-3}
for r in $(seq $R); do for L in $A $B; do
  PCG_LIB=$PWD/$L python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$L round $r kernel_ms=%.4f frac=%.3f'%(d['roofline']['kernel_ms'], d['roofline']['frac']))"
done; done
