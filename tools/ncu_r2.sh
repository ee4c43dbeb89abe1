#!/bin/bash
# ncu captures of the three dominant kernels of a build (plain run first, as the profiling recipe requires).
# usage: tools/ncu_r2.sh <tag>
T=${1:-r2x}; O=gpurun_out
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e"
$B > $O/plain_${T}_f32.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gik_solve_lane2 -c 1 -f -o $O/prof_${T}_f32 $B > $O/ncu_${T}_f32.log 2>&1
$B --config 4 > $O/plain_${T}_c4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gik_solve_pair -s 1 -c 1 -f -o $O/prof_${T}_edges $B --config 4 > $O/ncu_${T}_c4.log 2>&1
$B --dtype f64 > $O/plain_${T}_f64.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gik_solve_pair -c 1 -f -o $O/prof_${T}_f64 $B --dtype f64 > $O/ncu_${T}_f64.log 2>&1
ls -la $O/prof_${T}_*.ncu-rep
