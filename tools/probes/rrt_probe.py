"""Planner experiment (path_TESTS.py:910-990 in small) with the reference loop (expand=1) against the batched tree growth
(expand=8, 16): average time, iterations and success rate over the same seeded queries."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import gik_b200
from gik_b200 import experiments

solver = gik_b200.GraspIK(gik_b200.nextage_table(), "cuda:0").attach_scene()
a = (np.eye(3), np.array([0.33, -0.3, 0.93])); b = (np.eye(3), np.array([0.4, 0.11, 0.93]))
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 12
for K in (1, 8, 16):
    res = experiments.rrt_connect_trials(solver, a, b, std_devs=[(0.1, 0.1, 0.1), (0.2, 0.2, 0.2)], trials=trials, expand=K,
                                         generator=torch.Generator(device="cuda:0").manual_seed(4), rng=np.random.default_rng(4))
    for r in res:
        print(f"expand={K} sd={r['std_dev'][0]}: success {r['success_rate']:.0f} %, avg time {r['avg_time']:.3f} s (max {r['max_time']:.3f}), "
              f"avg iterations {r['avg_iterations']:.1f}", flush=True)
