#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python tools/summarize_ncu.py launches <launches.csv> <out.md> "<title>"
    python tools/summarize_ncu.py kernel   <report.ncu-rep> <out.md> "<title>"     (needs `ncu` on PATH)
"""
import csv
import io
import subprocess
import sys

UNIT_MS = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "sm__sass_thread_inst_executed_op_fadd_pred_on.sum", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "sm__sass_thread_inst_executed_op_dmul_pred_on.sum", "sm__sass_thread_inst_executed_op_dadd_pred_on.sum",
]
STALLS = "smsp__average_warps_issue_stalled_"


def launches(path, out, title):
    rows = list(csv.reader(open(path)))
    start = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    agg, order = {}, []
    for r in rows[start + 1:]:
        if len(r) < len(hdr):
            continue
        k = r[ix["Kernel Name"]]
        ms = float(r[ix["Metric Value"]].replace(",", "")) * UNIT_MS.get(r[ix["Metric Unit"]], 1.0)
        order.append((r[ix["ID"]], k, r[ix["Grid Size"]], r[ix["Block Size"]], ms))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` launch list "
                "(cold-cache, serialised launches: compare SHARES, not absolutes).\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:110]}` | {n} | {ms:.3f} | {100 * ms / tot:.2f}% |\n")
        f.write("\n## every launch, in order\n\n| id | kernel | grid | block | ms |\n|---|---|---|---|---:|\n")
        for o in order:
            f.write(f"| {o[0]} | `{o[1][:70]}` | {o[2]} | {o[3]} | {o[4]:.4f} |\n")


def kernel(rep, out, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# {title}\n\n`ncu --set full --clock-control none --import-source on` ({rep.split('/')[-1]}).\n")
        for vals in rows[2:]:
            d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
            f.write(f"\n## {d.get('Kernel Name', ('?', ''))[0]}  grid {d.get('Grid Size', ('?',''))[0]} block {d.get('Block Size', ('?',''))[0]}\n\n")
            f.write("| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in d and d[k][0] != "":
                    f.write(f"| {k} | {d[k][0]} | {d[k][1]} |\n")
            f.write("\nWarp-stall reasons (warps per issue-active cycle):\n\n| reason | value |\n|---|---:|\n")
            st = sorted(((h[len(STALLS):].replace("_per_issue_active.ratio", ""), float(v[0] or 0)) for h, v in d.items()
                         if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio")), key=lambda t: -t[1])
            for n, v in st:
                if v > 0.001:
                    f.write(f"| {n} | {v:.3f} |\n")
        # opcode mix from the source page
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        h = [i for i, r in enumerate(srows) if r and r[0] == "Address"]
        if h:
            hd = srows[h[0]]
            ix = {k: i for i, k in enumerate(hd)}
            mix = {}
            mx = 0
            for r in srows[h[0] + 1:]:
                if len(r) < len(hd):
                    continue
                ex = int(r[ix["Instructions Executed"]])
                mx = max(mx, ex)
                toks = [t for t in r[ix["Source"]].split() if not t.startswith("@")]
                op = toks[0].split(".")[0]
                mix[op] = mix.get(op, 0) + ex
            tot = sum(mix.values())
            f.write(f"\nExecuted warp-instruction mix (first kernel; {tot} warp instructions, "
                    f"{tot / max(mx, 1):.0f} per trip of the hottest instruction):\n\n| opcode | share | per hot-loop trip |\n|---|---:|---:|\n")
            for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:18]:
                f.write(f"| {k} | {100 * v / tot:.1f}% | {v / max(mx, 1):.1f} |\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](*sys.argv[2:5])
