"""Latency of ONE drop-in call (config 1: the reference's own use): computeqgrasppose(q0, CUBE_PLACEMENT) in fp64 (the
drop-in's default) and fp32, wall clock around the Python call, plus the kernel alone (CUDA events, one problem)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
import gik_b200

dev = torch.device("cuda:0")
solver = gik_b200.GraspIK(gik_b200.nextage_table(), dev)
cube = (np.eye(3), np.array([0.33, -0.3, 0.93]))
q0 = np.zeros(15)
for dtype in (torch.float64, torch.float32):
    for _ in range(5):
        q, ok = gik_b200.computeqgrasppose(solver, q0, None, cube, dtype=dtype, collision=None)
    t = time.perf_counter()
    for _ in range(50):
        q, ok = gik_b200.computeqgrasppose(solver, q0, None, cube, dtype=dtype, collision=None)
    wall = (time.perf_counter() - t) / 50
    P = torch.tensor(np.concatenate([cube[0].reshape(9), cube[1]])[:, None], dtype=dtype, device=dev)
    Q = torch.zeros((15, 1), dtype=dtype, device=dev)
    for _ in range(5):
        out = solver.solve_soa(Q, P)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        out = solver.solve_soa(Q, P)
    e1.record(); torch.cuda.synchronize()
    print(f"{str(dtype):14s} drop-in call {wall*1e3:.3f} ms wall (success={ok}, iterations {int(out[2][0])}) | "
          f"kernel + launch {e0.elapsed_time(e1)/50:.3f} ms for one problem")

# the full predicate on the device (collision term + keep-descending tail): the golden placement (740 iterations, free), and
# the worst case -- a converged-but-colliding placement that descends to the iteration cap and is then decided by the tail
solver.attach_scene()
sys.path.insert(0, 'tests')
from conftest import make_poses
P = make_poses(96, 77)
_, conv = solver.solve(torch.zeros(15), torch.as_tensor(P, device=dev), dtype=torch.float64)
_, succ, _, _, _ = solver.solve_success_soa(torch.zeros((15, 96), dtype=torch.float64, device=dev),
                                            torch.as_tensor(P, device=dev).t().contiguous())
bad = int(torch.nonzero(conv & ~succ.bool())[0])
for name, place in (("golden placement", cube), ("converged-but-colliding placement (worst case)", (np.eye(3), P[bad, 9:]))):
    for _ in range(5):
        q, ok = gik_b200.computeqgrasppose(None, q0, None, place)
    t = time.perf_counter()
    for _ in range(30):
        q, ok, it, _ = gik_b200.computeqgrasppose(None, q0, None, place, return_info=True)
    print(f"full predicate, fp64, {name}: {(time.perf_counter() - t) / 30 * 1e3:.3f} ms wall per drop-in call (success={ok}, iterations {it})")
