#!/bin/bash
# r2d: packed-log6 builds (MINB 3/4/5) on config 2 fp32, and a pairs-per-warp sweep of the edge projection
one() { python bench.py "$@" --steps 5 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('value %.3fM' % (d['value']/1e6), 'kernel_ms %.3f' % r['kernel_ms'], r['kernel'], 'conv %.4f' % d['config']['converged_fraction'])"; }
for lib in build/libgik_p6.so build/libgik_p6w4.so build/libgik_p6w5.so; do
  export GIK_LIB=$PWD/$lib; echo -n "$lib cfg2 f32: "; one --config 2 --dtype f32
done
export GIK_LIB=$PWD/build/libgik_p6.so
for n in 16384 65536 262144; do echo -n "p6 cfg2 f32 n=$n: "; one --config 2 --dtype f32 --n $n; done
unset GIK_LIB
for pw in 1 2 3 4 5 7 10 16; do
  echo -n "default lib cfg4 f32 GIK_PER_WARP=$pw: "; GIK_PER_WARP=$pw one --config 4 --dtype f32
done
for pw in 2 4 7 16; do
  echo -n "default lib cfg4 f64 GIK_PER_WARP=$pw: "; GIK_PER_WARP=$pw one --config 4 --dtype f64
done
for n in 16 256 1024; do echo -n "cfg4 f32 edges=$n: "; one --config 4 --dtype f32 --n $n; done
