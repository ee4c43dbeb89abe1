"""Single-GPU probe of the fused all-gather's push path: two result arrays on ONE device stand in for (self, peer), so
the pushers' loads / polls are measured without NVLink.  Run with GIK_FUSED_STATS=1 to print the per-warp statistics."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import gik_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
peers = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
solver = gik_b200.GraspIK(gik_b200.nextage_table(), dev)
g = torch.Generator(device=dev).manual_seed(5)
pos = torch.tensor([0.2, -0.4, 0.93], device=dev) + torch.rand((n, 3), device=dev, generator=g) * torch.tensor([0.4, 0.8, 0.47], device=dev)
pose = torch.cat([torch.eye(3, device=dev).reshape(1, 9).expand(n, 9), pos], 1).t().contiguous()
q0 = torch.zeros((15, n), device=dev)
qs = [torch.full((15, n), -1.0, device=dev) for _ in range(peers)]
cs = [torch.full((n,), 7, dtype=torch.uint8, device=dev) for _ in range(peers)]
for rep in range(3):
    torch.cuda.synchronize(); t = time.time()
    solver.solve_scatter_soa(q0, pose, [x.data_ptr() for x in qs], [x.data_ptr() for x in cs], n, 0)
    torch.cuda.synchronize(); print(f"call {rep}: {(time.time() - t) * 1e3:.3f} ms", flush=True)
q, c, _, _ = solver.solve_soa(q0, pose)
ok = all(torch.equal(q, x) for x in qs) and all(torch.equal(c, x) for x in cs)
print("equal to the plain solve:", ok)
