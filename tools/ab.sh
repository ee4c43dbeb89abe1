#!/bin/bash
# A/B the solver builds under build/*.so on the GPU box: tools/ab.sh [bench args]; prints solves/s and roofline per build
for so in build/libgik_*.so; do
  GIK_LIB=$PWD/$so python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); r = d['roofline']
print('$so', '%.3fM solves/s  kernel %.2f ms  roofline %.3f  conv %.4f' % (d['value']/1e6, r['kernel_ms'], r['frac'], d['config']['converged_fraction']))"
done
