#!/bin/bash
# A/B of library builds on one box: config 2 fp32 / fp64 and config 4, kernel-only figures (bench.py without the CPU / e2e legs).
# usage: tools/ab_r2.sh tag lib [lib ...]   (lib = path of a libgik build, "default" = the in-tree one)
tag=$1; shift
for lib in "$@"; do
  name=$(basename $lib .so)
  if [ "$lib" = "default" ]; then unset GIK_LIB; else export GIK_LIB=$PWD/$lib; fi
  for cfg in "2 f32" "2 f64" "4 f32" "4 f64"; do
    set -- $cfg
    python bench.py --config $1 --dtype $2 --steps 5 --warmup 2 --no-cpu-baseline --no-e2e $EXTRA 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$tag $name cfg$1 $2', 'value %.3fM' % (d['value']/1e6), 'kernel_ms %.3f' % r['kernel_ms'], r['kernel'], 'frac %.3f' % r['frac'], 'conv %.4f' % d['config']['converged_fraction'])"
  done
done
