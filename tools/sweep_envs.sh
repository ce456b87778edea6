#!/bin/bash
# kernel time vs batch size for the monolithic (PCG_WS=0) and warp-specialised (PCG_WS=1|2) kernels (profiling aid)
for n in ${SWEEP_N:-9472 18944 37888 65536 131072 262144}; do for w in ${SWEEP_WS:-0 1 3}; do for nm in ${SWEEP_NOISE:-off philox}; do
  PCG_WS=$w python bench.py --steps 10 --warmup 3 --no-cpu-baseline --n-envs $n --noise-mode $nm 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); ms=d['roofline']['kernel_ms']; print('n=$n ws=$w noise=$nm kernel_ms=%.4f cycles/interval=%.0f value=%.3e frac=%.3f'%(ms, ms*1e-3*1.965e9/99, d['value'], d['roofline']['frac']))"
done; done; done
