#!/usr/bin/env python
"""bench.py -- dual-arm grasp IK solves/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--dtype f32|f64]

A "step" is ONE pass of the hot path over one batch of synthetic input: BASELINE config 2, 2^20 random cube
placements over the table workspace (identity rotation, path.py:47), every problem started from q0 = 0,
reference preset (eps 1e-3, dt 1e-2, max_iters 1000, undamped), fp32.  For N > 1 every rank solves its own
2^20-problem slab (weak scaling, no data-path collective) and the step ends with the all-gather of (q, converged).

Printed JSON (one line, rank 0): value = whole-job solves/s with inputs resident in HBM; e2e = the same metric
through the public API with pinned HOST buffers (H2D + solve + D2H inside the timed region); roofline = the solve
kernel against the CUDA-core FMA peak measured in the same run; cpu_baseline = the C oracle (a port of the
reference's loop; the reference itself needs pinocchio, absent from this image) on the box's host cores.
`--impl reference` times that CPU port alone."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PER_GPU = 1 << 20
WORKSPACE_LO = (0.20, -0.40, 0.93)
WORKSPACE_HI = (0.60, 0.40, 1.40)
SAMPLER_LO = (0.33, -0.30, 1.05)        # the reference sampler's box, path.py:35-37 with config.py:36-37
SAMPLER_HI = (0.40, 0.11, 1.40)
METRIC = "dual-arm grasp IK solves/sec"
_STDOUT = sys.stdout


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default=None, choices=["f32", "f64"], help="default: f32 (f64 for --config 3)")
    ap.add_argument("--n", type=int, default=N_PER_GPU, help="problems per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=0, help="problems in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json config: 2 (default, the metric's workload), 3 restarts fp64, 4 edge projection, 5 sharded sweep")
    ap.add_argument("--no-gather", action="store_true", help="N > 1: skip the result all-gather")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"],
                    help="N > 1: 'fused' = the solve kernel stores results into every rank's symmetric-memory arrays over "
                         "NVLink (falls back to nccl if symmetric memory is unavailable); 'nccl' = all_gather_into_tensor")
    ap.add_argument("--box", default="workspace", choices=["workspace", "sampler"],
                    help="config 2 placement box: the table workspace (default) or the reference sampler's box (path.py:35-37)")
    ap.add_argument("--early-stop", action="store_true",
                    help="fast preset GIK_F_EARLY_STOP (NOT the reference's semantics for failed problems; never the default)")
    ap.add_argument("--kernel", default=None, choices=["lane", "pair", "lane1"], help="force a thread mapping (default: launcher's choice)")
    ap.add_argument("--step", default="auto", choices=["auto", "cholesky"],
                    help="form of the undamped step: auto = spherical-wrist 3x3 solves where the table allows it (default), "
                         "cholesky = always the general 6x6 block-Cholesky form (A/B)")
    args = ap.parse_args()
    if args.dtype is None:
        args.dtype = "f64" if args.config == 3 else "f32"
    return args


def workload_name(n, dtype):
    return (f"config2: {n} random cube placements over the table workspace x[0.20,0.60] y[-0.40,0.40] z[0.93,1.40], "
            f"identity rotation, single start q0=0, {dtype}, eps=1e-3 dt=1e-2 max_iters=1000 damping=0")


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (bounded sample of the same workload)
# ----------------------------------------------------------------------------------------------------------
def host_poses(n, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    P = np.zeros((n, 12))
    P[:, [0, 4, 8]] = 1.0
    P[:, 9:] = rng.uniform(WORKSPACE_LO, WORKSPACE_HI, size=(n, 3))
    return P


def cpu_solve_rate(n_sample, seed=1234):
    """Times the C oracle (oracle/grasp_ik_oracle.c, OpenMP over problems, all host threads) on n_sample problems
    of the bench workload.  Returns (solves/s, threads, seconds)."""
    import numpy as np
    import gik_b200
    from oracle import c_oracle
    c_oracle.build()
    tc = gik_b200.nextage_table().to_c()
    threads = os.cpu_count() or 1
    P = host_poses(n_sample, seed)
    c_oracle.solve(tc, np.zeros((min(threads, n_sample), 15)), P[:min(threads, n_sample)])   # warm the threads
    t0 = time.perf_counter()
    c_oracle.solve(tc, np.zeros((n_sample, 15)), P, threads=threads)
    dt = time.perf_counter() - t0
    return n_sample / dt, threads, dt


def numpy_solve_rate(n_sample, seed=99):
    """The numpy restatement (oracle/grasp_ik_np.py: one Python-level call per pinocchio/numpy primitive, like the
    reference's own loop) on one core -- the closest thing to the reference's Python speed this image can run."""
    import numpy as np
    from oracle import grasp_ik_np as o
    P = host_poses(n_sample, seed)
    t0 = time.perf_counter()
    for i in range(n_sample):
        o.computeqgrasppose(np.zeros(15), np.eye(3), P[i, 9:])
    return n_sample / (time.perf_counter() - t0)


def run_reference(args):
    """`--impl reference`: the reference's own CPU algorithm for the path.  The reference is Python on pinocchio
    (not installable here: no wheel, no network), so the arm times the C port of its loop (oracle/), which
    reproduces the reference's golden outputs; every step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_sample = args.cpu_sample or 16 * cores
    for _ in range(max(args.warmup, 0) and 1):
        cpu_solve_rate(min(n_sample, 2 * cores))
    t_total = 0.0
    for k in range(args.steps):
        rate, threads, dt = cpu_solve_rate(n_sample, seed=1234 + k)
        t_total += dt
    value = n_sample * args.steps / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(N_PER_GPU, "fp64 on CPU"),
                   "sample_per_step": n_sample},
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": threads, "kind": "port",
                         "sample": f"{n_sample} problems of the workload per step x {args.steps} steps, C oracle "
                                   f"(Jacobi-SVD pinv, -O3 -march=native, OpenMP {threads} threads)"},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_STDOUT, flush=True)


# ----------------------------------------------------------------------------------------------------------
# clocks during the timed region
# ----------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.sm_max = None
        self._halt = threading.Event()
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for line in self.proc.stdout:
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    self.samples.append(float(f[0])); self.sm_max = float(f[1])
                except ValueError:
                    continue
                for nm, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
                if self._halt.is_set():
                    break
        except Exception:
            pass

    def stop(self):
        self._halt.set()
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
def workspace_positions(torch, n, dev, dtype, seed, box="workspace"):
    g = torch.Generator(device=dev).manual_seed(seed)
    lo = torch.tensor(WORKSPACE_LO if box == "workspace" else SAMPLER_LO, device=dev, dtype=dtype)
    hi = torch.tensor(WORKSPACE_HI if box == "workspace" else SAMPLER_HI, device=dev, dtype=dtype)
    return lo + torch.rand((n, 3), device=dev, dtype=dtype, generator=g) * (hi - lo)


def pose_rows_from(torch, pos):
    n = pos.shape[0]
    return torch.cat([torch.eye(3, device=pos.device, dtype=pos.dtype).reshape(1, 9).expand(n, 9), pos], 1).contiguous()


class Workload:
    """One BASELINE config on one rank: `launch()` enqueues the kernels of one step on the current stream and
    returns the tensors a gather would ship; `solves` = problems solved per step on this rank; `iterations()` = descent
    iterations executed by the last step (device scalar); `converged()` = fraction converged."""
    name = ""
    solves = 0
    kernel_choice = None
    early_stop = False


class Config2(Workload):
    def __init__(self, torch, solver, n, rank, dtype, seed_base=1000, box="workspace"):
        dev = solver.device
        self.torch, self.solver, self.solves = torch, solver, n
        self.pose_rows = pose_rows_from(torch, workspace_positions(torch, n, dev, dtype, seed_base + rank, box))
        self.pose = self.pose_rows.t().contiguous()                 # SoA [12][n]
        self.q0 = torch.zeros((15, n), device=dev, dtype=dtype)     # SoA [15][n]
        self.out = (torch.empty_like(self.q0), torch.empty(n, dtype=torch.uint8, device=dev),
                    torch.empty(n, dtype=torch.int32, device=dev), torch.empty((2, n), dtype=dtype, device=dev))
        self.name = workload_name(n, "fp32" if dtype == torch.float32 else "fp64")
        if box == "sampler":
            self.name = self.name.replace("over the table workspace x[0.20,0.60] y[-0.40,0.40] z[0.93,1.40]",
                                          "over the reference sampler box x[0.33,0.40] y[-0.30,0.11] z[1.05,1.40] (path.py:35-37)")

    def launch(self):
        q, conv, _, _ = self.solver.solve_soa(self.q0, self.pose, out=self.out, kernel=self.kernel_choice,
                                              early_stop=self.early_stop)
        return q, conv

    def iterations(self):
        return self.out[2].sum(dtype=self.torch.int64).double()

    def converged(self):
        return self.out[1].double().mean()


class Config3(Workload):
    """65,536 placements x 64 random-restart initial q (restart 0 = zeros), best-of, fp64."""
    DAMPING = 1e-6

    def __init__(self, torch, solver, n_place, restarts, rank, dtype):
        dev = solver.device
        self.torch, self.solver, self.n_place, self.R = torch, solver, n_place, restarts
        self.solves = n_place * restarts
        pos = workspace_positions(torch, n_place, dev, dtype, 1 + 7919 * rank)
        rows = pose_rows_from(torch, pos)
        self.pose = rows.unsqueeze(1).expand(n_place, restarts, 12).reshape(-1, 12).t().contiguous()
        lo, hi = solver.limits(dtype)
        g = torch.Generator(device=dev).manual_seed(2 + 7919 * rank)
        cand = lo + torch.rand((n_place, restarts, 15), device=dev, dtype=dtype, generator=g) * (hi - lo)
        cand[:, 0, :] = 0
        self.q0 = cand.reshape(-1, 15).t().contiguous()
        n = self.solves
        self.out = (torch.empty_like(self.q0), torch.empty(n, dtype=torch.uint8, device=dev),
                    torch.empty(n, dtype=torch.int32, device=dev), torch.empty((2, n), dtype=dtype, device=dev))
        self.best = None
        self.name = (f"config3: {n_place} workspace placements x {restarts} restarts (restart 0 = q0, others uniform in the "
                     f"joint limits), best-of, {'fp64' if dtype == torch.float64 else 'fp32'}, damping {self.DAMPING}")

    def launch(self):
        q, conv, _, resid = self.solver.solve_soa(self.q0, self.pose, out=self.out, damping=self.DAMPING, kernel=self.kernel_choice)
        self.best = self.solver.best_of_soa(q, conv, resid, self.n_place, self.R)
        return self.best[0], self.best[1]

    def iterations(self):
        return self.out[2].sum(dtype=self.torch.int64).double()

    def converged(self):
        return self.out[1].double().mean()

    def extra(self):
        return {"placements_with_a_converged_restart": float(self.best[1].double().mean()),
                "restart0_converged": float(self.out[1].reshape(self.n_place, self.R)[:, 0].double().mean())}


class Config4(Workload):
    """path.py edge projection: E edges x S interpolated placements, warm-started along each edge, march stops at the
    first non-converged step.  Step 0 starts from a converged q at cube_a."""

    def __init__(self, torch, solver, n_edges, steps, rank, dtype):
        dev = solver.device
        self.torch, self.solver, self.E, self.S = torch, solver, n_edges, steps
        # endpoints a: workspace placements whose IK from q0 converges (oversample, keep the first E)
        cand = pose_rows_from(torch, workspace_positions(torch, 3 * n_edges, dev, dtype, 3 + 7919 * rank))
        q, conv = solver.solve(torch.zeros(15, device=dev, dtype=dtype), cand, dtype=dtype)
        keep = torch.nonzero(conv).flatten()[:n_edges]
        assert keep.numel() == n_edges, "not enough reachable edge starts"
        self.q_start = q[keep].t().contiguous()
        self.pose_a = cand[keep].t().contiguous()
        self.pose_b = pose_rows_from(torch, workspace_positions(torch, n_edges, dev, dtype, 4 + 7919 * rank)).t().contiguous()
        self.ns = torch.full((n_edges,), steps, dtype=torch.int32, device=dev)
        self.res = None
        self.solves = n_edges * steps          # upper bound; the executed count is reported from n_valid
        self.name = (f"config4: {n_edges} cube-pose edges x {steps} interpolated placements (SE3.Interpolate), IK "
                     f"warm-started along each edge, stop at first failure, {'fp32' if dtype == torch.float32 else 'fp64'}")

    def launch(self):
        self.res = self.solver.project_edges_soa(self.q_start, self.pose_a, self.pose_b, self.ns, self.S, kernel=self.kernel_choice)
        return self.res[0], self.res[1]

    def executed(self):
        nv = self.res[1]
        return (nv + (nv < self.S).to(nv.dtype)).sum(dtype=self.torch.int64).double()   # + the failing solve

    def iterations(self):
        return self.res[2].sum(dtype=self.torch.int64).double()

    def converged(self):
        return self.res[1].double().sum() / self.executed()

    def extra(self):
        nv = self.res[1].double()
        return {"edges_fully_projected": float((nv == self.S).double().mean()), "mean_valid_steps": float(nv.mean()),
                "solves_executed": float(self.executed())}


def run_b200(args):
    import torch
    import torch.distributed as dist
    import gik_b200
    from gik_b200 import dist as gdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    esz = 4 if args.dtype == "f32" else 8
    solver = gik_b200.GraspIK(gik_b200.nextage_table(), dev)
    solver.force_cholesky = args.step == "cholesky"

    scaling = "weak"
    if args.config == 2:
        wl = Config2(torch, solver, args.n, rank, dtype, box=args.box)
    elif args.config == 3:
        wl = Config3(torch, solver, args.n if args.n != N_PER_GPU else 65536, 64, rank, dtype)
    elif args.config == 4:
        wl = Config4(torch, solver, args.n if args.n != N_PER_GPU else 4096, 256, rank, dtype)
    else:   # config 5: 64 Mi config-2 problems sharded over the ranks (strong scaling) + all-gather
        total = args.n if args.n != N_PER_GPU else 64 << 20
        lo_i, hi_i = gdist.shard_bounds(total, rank, world)
        wl = Config2(torch, solver, hi_i - lo_i, rank, dtype, seed_base=4000)
        wl.name = f"config5: {total} config-2 problems sharded over {world} GPU(s) + all-gather of q/converged; " + wl.name
        scaling = "strong"
    wl.kernel_choice = args.kernel
    wl.early_stop = args.early_stop
    if args.early_stop:
        wl.name += " [FAST PRESET: early stop of stalled problems -- not the reference's semantics for failed problems]"
    gather = world > 1 and not args.no_gather and args.config in (2, 5)
    if args.config == 5:
        n_total_gather, gather_off = total, lo_i
    else:
        n_total_gather, gather_off = wl.solves * world, wl.solves * rank
    fused = None
    collective = "none"
    if gather:
        collective = "all_gather_into_tensor (NCCL) of q [15][n] + converged [n] per step"
        if args.gather == "fused":
            try:
                fused = gdist.SymmetricResults(15, n_total_gather, dtype, dev)
                fused_out = (wl.out[2], wl.out[3])
                collective = ("fused: the solve kernel's epilogue stores q/converged into every rank's symmetric-memory "
                              "result arrays over NVLink (P2P st.global), then one cross-rank barrier per step")
            except Exception as e:      # symmetric memory unavailable on this box: keep the NCCL gather and say so
                fused = None
                collective += f" (fused path unavailable: {type(e).__name__})"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(ev=None):
        flush.fill_(1)                                       # L2 flush between iterations
        if ev:
            ev[0].record()
        if fused is not None:
            solver.solve_scatter_soa(wl.q0, wl.pose, fused.q_ptrs, fused.conv_ptrs, n_total_gather,
                                     gather_off, out=fused_out, kernel=args.kernel)
        else:
            q, conv = wl.launch()
        if ev:
            ev[1].record()
        if fused is not None:
            fused.barrier()
        elif gather:
            gdist.all_gather_results(q, conv, n_total_gather)

    for _ in range(max(args.warmup, 0)):
        step()
    barrier()

    # kernel-only events (on the launching stream = torch's current stream) for the roofline
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = solver.launches
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        step(kev[k])
    e1.record()
    barrier()
    if sampler:
        sampler.stop()
    if fused is not None:     # the local flags live in the symmetric array: bring this rank's slab back for the stats
        wl.out[1].copy_(fused.conv[gather_off:gather_off + wl.solves])
    ms_total = e0.elapsed_time(e1)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / max(args.steps, 1)
    launches = solver.launches - launches0
    t = torch.tensor([ms_total, kernel_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, kernel_ms = t.tolist()
    ms_per_step = ms_total / max(args.steps, 1)

    solves_step = wl.executed() if hasattr(wl, "executed") else torch.tensor(float(wl.solves), device=dev, dtype=torch.float64)
    stats = torch.stack([wl.converged() * solves_step, wl.iterations(), solves_step])
    if world > 1:
        dist.all_reduce(stats)
    conv_solves, iters_sum, solves_all = stats.tolist()          # per step, all ranks
    conv_frac = conv_solves / solves_all

    # ---- e2e: public batched API with pinned HOST buffers, copies inside the timed region (default workload only)
    e2e = None
    if not args.no_e2e and args.config == 2:
        n = wl.solves
        q_host = torch.zeros((n, 15), dtype=dtype).pin_memory()
        pose_host = wl.pose_rows.cpu().pin_memory()

        def e2e_step():
            # public API, HOST tensors in and out: H2D of q_init/pose, the solve and D2H of q/converged all happen
            # inside this call (pipelined over slabs by GraspIK.solve_host); it returns when the results are in host memory
            qq, cc = gik_b200.computeqgrasppose_batch(solver, q_host, pose_host, dtype=dtype)
            return bool(cc[0])

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(args.steps):
            e2e_step()
        a1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        te = torch.tensor([max(a0.elapsed_time(a1), wall)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": n * world * args.steps / (te.item() * 1e-3), "unit": "solves/s",
               "h2d_bytes_per_step": int(n * (15 + 12) * esz), "d2h_bytes_per_step": int(n * (15 * esz + 1)),
               "api": "computeqgrasppose_batch(host tensors) -> host tensors (pinned row-major in/out; one launch, inputs streamed by the copy engine, results stored to host memory by the kernel)"}

    extra = wl.extra() if hasattr(wl, "extra") else {}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the solve kernel the launcher picked): CUDA-core FMA pipe, peak measured in this run
    fpi = gik_b200.flops_per_iter()
    peak = gik_b200.fma_peak_tflops(local, esz)
    iters_per_launch = iters_sum / world
    achieved = iters_per_launch * fpi / (kernel_ms * 1e-3) * 1e-12
    bytes_algo = gik_b200.bytes_per_solve(esz) * solves_all / world
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    traffic = pipe_busy = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get(args.dtype) if args.config == 2 else None
        pipe_busy = tj.get("pipe_busy", {}).get(args.dtype) if args.config == 2 else None
    except Exception:
        pass
    roofline = {
        "bound": "fp32" if esz == 4 else "fp64",
        "kernel": solver.kernel_name(wl.E if args.config == 4 else wl.solves, dtype, args.kernel)
        + (" (edge mode)" if args.config == 4 else ""),
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
        "peak_source": "measured in this run: register-resident FMA chains on every SM (gik_measure_fma_peak); "
                       "MEASURED_PEAKS.json has no CUDA-core figure",
        "flops_per_iteration": fpi, "iterations_per_launch": iters_per_launch, "kernel_ms": kernel_ms,
        "traffic": traffic,
        # `frac` uses the survey's ALGORITHMIC FLOP count (3117 / iteration); the kernel executes fewer (DESIGN.md
        # "Roofline"), so frac can exceed 1.  How busy the pipe really is comes from the committed ncu capture:
        "pipe_busy_ncu": pipe_busy,
        "hbm": {"algorithmic_bytes_per_launch": bytes_algo,
                "achieved_gbs": bytes_algo / (kernel_ms * 1e-3) * 1e-9, "peak_gbs": hbm_peak,
                "frac": (bytes_algo / (kernel_ms * 1e-3) * 1e-9 / hbm_peak) if hbm_peak else None},
    }

    cpu = None
    if not args.no_cpu_baseline and world == 1 and args.config == 2:
        cores = os.cpu_count() or 1
        n_sample = args.cpu_sample or 512 * cores
        rate, threads, secs = cpu_solve_rate(n_sample)
        np_rate = numpy_solve_rate(3)
        cpu = {"value": rate, "unit": "solves/s", "cores": threads, "kind": "port",
               "numpy_restatement_solves_per_s_1core": np_rate,
               "sample": f"{n_sample} problems of the same workload ({secs:.1f} s), C oracle (Jacobi-SVD pinv, -O3 "
                         f"-march=native, OpenMP {threads} threads); the Python+pinocchio reference cannot be installed "
                         f"here -- its closest stand-in, the numpy restatement, is timed on 3 problems / 1 core beside it"}

    cfg = {"workload": wl.name, "problems_per_gpu_per_step": solves_all / world,
           "l2": "256 MB flush write before every step (inside the timed region)",
           "collective": collective,
           "converged_fraction": conv_frac, "mean_iterations": iters_sum / solves_all}
    cfg.update(extra)
    line = {
        "metric": METRIC, "value": solves_all * args.steps / (ms_total * 1e-3), "unit": "solves/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": cfg,
        "converged_solves_per_s": conv_solves * args.steps / (ms_total * 1e-3),
        "clocks": sampler.summary() if sampler else None,
        "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line), file=_STDOUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    # Exactly ONE line on stdout: libraries (NCCL prints its version banner there) are routed to stderr for the whole
    # run and the JSON line goes to the saved descriptor.
    global _STDOUT
    sys.stdout.flush()
    _STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
