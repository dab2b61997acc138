#!/bin/bash
# tuning helper: time the pCN kernel for every library build under bridge.jl_b200/lib/var/ (run under gpurun)
for lib in bridge.jl_b200/lib/var/*.so; do
  for chains in ${CHAINS:-250000}; do
    BB_LIB=$PWD/$lib timeout 300 python bench.py --steps ${STEPS:-10} --warmup 3 --no-e2e --no-cpu --chains $chains 2>&1 | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$lib', $chains, 'ms=%.3f'%r['kernel_ms'], 'GB/s=%.0f'%r['achieved'], 'frac=%.3f'%r['frac'], 'steps/s=%.3e'%d['value'], d['clocks']['sm_mhz'], d['clocks']['power_w_max'])"
  done
done
