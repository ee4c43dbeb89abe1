#!/usr/bin/env python
"""Throughput of the collision kernels and of the full success pipeline (GPU box): python tools/bench_collision.py"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gik_b200  # noqa: E402


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dev = torch.device("cuda:0")
    s = gik_b200.GraspIK(gik_b200.nextage_table(), dev).attach_scene()
    n = 1 << 20
    g = torch.Generator(device=dev).manual_seed(0)
    lo = torch.tensor([0.20, -0.40, 0.93], device=dev); hi = torch.tensor([0.60, 0.40, 1.40], device=dev)
    pos = lo + torch.rand((n, 3), device=dev, generator=g) * (hi - lo)
    pose = torch.cat([torch.eye(3, device=dev).reshape(1, 9).expand(n, 9), pos], 1).t().contiguous()
    q0 = torch.zeros((15, n), device=dev)
    q, conv, _, _ = s.solve_soa(q0, pose)
    out = {}
    col = s.collision_soa(q, pose).bool()
    out["converged_frac"] = conv.float().mean().item()
    out["colliding_frac_of_converged"] = (col & conv.bool()).float().sum().item() / conv.float().sum().item()
    t_solve = timed(lambda: s.solve_soa(q0, pose), 3)
    t_succ = timed(lambda: s.solve_success_soa(q0, pose, descend_while_colliding=False), 3)
    out["solve_ms"] = t_solve
    out["solve_plus_collision_ms"] = t_succ
    out["success_solves_per_s_M"] = n / t_succ / 1e3
    for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
        qq, pp = q.to(dt), pose.to(dt)
        out[f"collision_{name}_Mcfg_s"] = n / timed(lambda: s.collision_soa(qq, pp)) / 1e3
        out[f"clearance_{name}_Mcfg_s"] = n / timed(lambda: s.clearance_soa(qq, pp, 0.04)) / 1e3
        out[f"cube_collision_{name}_Mcfg_s"] = n / timed(lambda: s.cube_collision_soa(pp)) / 1e3
    t0 = time.perf_counter()
    a = (torch.eye(3), torch.tensor([0.33, -0.3, 0.93])); b = (torch.eye(3), torch.tensor([0.4, 0.11, 0.93]))
    import numpy as np
    qs, pl, ok = gik_b200.sample_grasp_poses_batch(s, 1 << 18, (np.eye(3), np.array([0.33, -0.3, 0.93])),
                                                   (np.eye(3), np.array([0.4, 0.11, 0.93])))
    torch.cuda.synchronize()
    out["sampler_262144_candidates_ms"] = (time.perf_counter() - t0) * 1e3
    out["sampler_accept_rate"] = ok.float().mean().item()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
