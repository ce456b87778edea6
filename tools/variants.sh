#!/bin/bash
# kernel-time matrix for the headline workload: envs per thread x noise mode (profiling aid, not a bench value)
for ept in 2 1; do for nm in philox philox_f32 off; do
  PCG_EPT=$ept python bench.py --steps 10 --warmup 3 --no-cpu-baseline --noise-mode $nm 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ept=$ept noise=$nm kernel_ms=%.4f value=%.3e frac=%.3f'%(d['roofline']['kernel_ms'], d['value'], d['roofline']['frac']))"
done; done
