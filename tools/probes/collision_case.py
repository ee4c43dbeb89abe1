"""One collision-kernel workload for ncu: 2^18 converged-ish configurations against the reference scene (fp32)."""
import sys, torch
sys.path.insert(0, '.')
import gik_b200
dev = torch.device("cuda:0")
s = gik_b200.GraspIK(gik_b200.nextage_table(), dev).attach_scene()
n = 1 << 18
g = torch.Generator(device=dev).manual_seed(0)
lo = torch.tensor([0.20, -0.40, 0.93], device=dev); hi = torch.tensor([0.60, 0.40, 1.40], device=dev)
pos = lo + torch.rand((n, 3), device=dev, generator=g) * (hi - lo)
pose = torch.cat([torch.eye(3, device=dev).reshape(1, 9).expand(n, 9), pos], 1).t().contiguous()
q, conv, _, _ = s.solve_soa(torch.zeros((15, n), device=dev), pose, early_stop=True)
for _ in range(3):
    c = s.collision_soa(q, pose)
torch.cuda.synchronize()
print("colliding fraction", c.float().mean().item())
