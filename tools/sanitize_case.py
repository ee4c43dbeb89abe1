#!/usr/bin/env python
"""Small pass through every kernel of libgik.so, for `compute-sanitizer --tool memcheck python tools/sanitize_case.py`
(one sanitizer tool per gpurun call; sizes kept tiny: the tool slows kernels down 10-50x)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gik_b200  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    s = gik_b200.GraspIK(gik_b200.nextage_table(), dev).attach_scene()
    rng = np.random.default_rng(0)
    for dtype in (torch.float32, torch.float64):
        for n in (1, 33, 257):
            P = torch.zeros((n, 12), dtype=dtype, device=dev); P[:, [0, 4, 8]] = 1
            P[:, 9:] = torch.as_tensor(rng.uniform([0.2, -0.4, 0.93], [0.6, 0.4, 1.4], size=(n, 3)), dtype=dtype, device=dev)
            pose = P.t().contiguous(); q0 = torch.zeros((15, n), dtype=dtype, device=dev)
            for kern in ("lane", "lane1", "pair"):
                q, conv, it, res = s.solve_soa(q0, pose, max_iters=40, kernel=kern)
            q, conv, it, res = s.solve_soa(q0, pose, max_iters=200, early_stop=True)
            s.fk_soa(q); s.jac_soa(q)
            s.collision_soa(q, pose); s.clearance_soa(q, pose, 0.04); s.cube_collision_soa(pose)
            s.best_of_soa(q, conv, res, 1, n)
            ns = torch.full((n,), 3, dtype=torch.int32, device=dev)
            for kern in ("lane", "pair"):
                s.project_edges_soa(q0, pose, pose + 0.01, ns, 3, max_iters=60, kernel=kern)
            qa = [torch.zeros((15, 2 * n + 5), dtype=dtype, device=dev) for _ in range(2)]
            ca = [torch.zeros((2 * n + 5,), dtype=torch.uint8, device=dev) for _ in range(2)]
            s.solve_scatter_soa(q0, pose, [t.data_ptr() for t in qa], [t.data_ptr() for t in ca], 2 * n + 5, n, max_iters=30)
            s.solve_host(torch.zeros(15), P.cpu(), dtype=dtype, max_iters=30)
            s.solve_success_soa(q0, pose, max_iters=120)
    torch.cuda.synchronize()
    print("sanitize_case: done")


if __name__ == "__main__":
    main()
