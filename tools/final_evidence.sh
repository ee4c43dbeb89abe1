#!/bin/bash
# One GPU-box pass that refreshes the evidence under gpurun_out/ for a build: smoke, GPU tests (which rewrite
# profiles/parity_r2.json), the default bench line with all sub-runs, probes, the ncu launch list of the default bench
# command and `ncu --set full` captures of the three dominant kernels.   usage: tools/final_evidence.sh <tag>   (e.g. r2s)
# Multi-GPU lines (run under `gpurun --gpus N`):
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
#       bench.py --gpus N --steps 10 --warmup 3 --no-cpu-baseline [--gather nccl | --no-gather]
T=${1:-rX}; O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$T.log 2>&1; tail -1 $O/smoke_$T.log
python -m pytest tests -m gpu -q > $O/gpu_tests_$T.log 2>&1; tail -2 $O/gpu_tests_$T.log
cp profiles/parity_r2.json $O/parity_r2_$T.json
python bench.py > $O/bench_${T}_cfg2.json 2> $O/b.err
python tools/probes/single_solve_latency.py > $O/single_latency_$T.log 2>&1
python tools/probes/rrt_probe.py 12 > $O/rrt_probe_$T.log 2>&1
python tools/probes/edge_agreement_probe.py > $O/edge_agreement_$T.log 2>&1
GIK_FUSED_STATS=1 python tools/probes/fused_probe.py 1048576 2 > $O/fused_probe_$T.log 2>&1
bash tools/ncu_r2.sh $T > $O/ncu_$T.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$T.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launches_$T.log 2>&1
cat $O/single_latency_$T.log $O/rrt_probe_$T.log; tail -12 $O/fused_probe_$T.log
python - <<EOF
import json
d = json.load(open("$O/bench_${T}_cfg2.json"))
print("config 2 fp32: %.2f M solves/s, %.2f ms, e2e %.2f M, frac_executed %.3f" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["roofline"]["frac_executed"]))
for k, v in d["sub"].items():
    if isinstance(v, dict) and "value" in v:
        print(k, "%.2f M solves/s, %.2f ms" % (v["value"] / 1e6, v["ms_per_step"]))
EOF
# then, here:  python tools/summarize_ncu.py kernel gpurun_out/prof_${T}_f32.ncu-rep profiles/${T}_solve_f32_ncu.md "<title>"  (f64, edges alike)
