#!/bin/bash
# kernel time vs batch size: monolithic (ws=0), pair-per-CTA (ws=1), per-SM finisher pool (ws=3, falls back where it is
# not eligible) and the library's own choice (ws=auto)   (profiling aid)
for n in ${SWEEP_N:-9472 18944 37888 65536 131072 262144}; do for w in ${SWEEP_WS:-0 1 3}; do for nm in ${SWEEP_NOISE:-off philox}; do
  if [ "$w" = "auto" ]; then unset PCG_WS; else export PCG_WS=$w; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --n-envs $n --noise-mode $nm 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); ms=d['roofline']['kernel_ms']; print('n=$n ws=$w noise=$nm kernel_ms=%.4f cycles/interval=%.0f value=%.3e frac=%.3f %s'%(ms, ms*1e-3*1.965e9/99, d['value'], d['roofline']['frac'], d['roofline']['kernel'].split(' ')[0][5:]))"
done; done; done
