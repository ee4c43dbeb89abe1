import sys, torch
sys.path.insert(0, '.')
import gik_b200
dev = torch.device("cuda:0")
s = gik_b200.GraspIK(gik_b200.nextage_table(), dev)
n = 1 << 20
g = torch.Generator(device=dev).manual_seed(12)
lo = torch.tensor([0.20, -0.40, 0.93], device=dev); hi = torch.tensor([0.60, 0.40, 1.40], device=dev)
pos = lo + torch.rand((n, 3), device=dev, generator=g) * (hi - lo)
pose = torch.cat([torch.eye(3, device=dev).reshape(1, 9).expand(n, 9), pos], 1).t().contiguous()
q0 = torch.zeros((15, n), device=dev)
for dtype in (torch.float32, torch.float64):
    a = s.solve_soa(q0.to(dtype), pose.to(dtype))
    b = s.solve_soa(q0.to(dtype), pose.to(dtype), early_stop=True)
    fa, fb = a[1].bool(), b[1].bool()
    fn = fa & ~fb
    print(dtype, "converged", fa.sum().item(), "false negatives", fn.sum().item(), "false positives", (fb & ~fa).sum().item(),
          "iters of FN (full run):", a[2][fn][:20].tolist(), "stopped at:", b[2][fn][:20].tolist(),
          "iterations saved", 1 - b[2].sum().item() / a[2].sum().item())
    print(" FN positions:", pos[fn][:8].tolist())
