#!/bin/bash
# One GPU-box pass that refreshes the evidence under gpurun_out/ for a build: smoke, GPU tests, bench lines of every
# config, probes, the ncu launch list of the default bench command and `ncu --set full` captures of the packed fp32
# lane kernel and of the edge-mode pair kernel.   usage: tools/final_evidence.sh <tag>     (e.g. r1k)
T=${1:-rX}; O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$T.log 2>&1
python -m pytest tests -m gpu -x -q > $O/gpu_tests_$T.log 2>&1; tail -2 $O/gpu_tests_$T.log
python bench.py > $O/bench_${T}_cfg2.json 2> $O/b.err
python bench.py --config 3 --steps 3 > $O/bench_${T}_cfg3.json 2>> $O/b.err
python bench.py --config 4 > $O/bench_${T}_cfg4.json 2>> $O/b.err
python bench.py --dtype f64 --steps 5 --no-cpu-baseline > $O/bench_${T}_cfg2_f64.json 2>> $O/b.err
python bench.py --config 4 --dtype f64 --steps 3 > $O/bench_${T}_cfg4_f64.json 2>> $O/b.err
python tools/probes/precision_probe.py > $O/precision_$T.log 2>&1
python tools/probes/single_solve_latency.py > $O/single_latency_$T.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$T.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gik_solve_lane2 -c 1 -o $O/prof_${T}_f32 -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $O/ncu_f32.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gik_solve_pair -s 1 -c 1 -o $O/prof_${T}_edges -f \
    python bench.py --config 4 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $O/ncu_c4.log 2>&1
cat $O/smoke_$T.log $O/precision_$T.log $O/single_latency_$T.log
for f in $O/bench_${T}_*.json; do python -c "
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['value']/1e6,3), d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel'], d['e2e'] and d['e2e']['value'])" $f; done
