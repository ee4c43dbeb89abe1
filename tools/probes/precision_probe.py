"""fp32 vs fp64 kernels on 2^20 workspace problems: flag agreement, |dq| quantiles, iteration differences."""
import sys, torch
sys.path.insert(0, '.')
import gik_b200
dev = torch.device("cuda:0")
s = gik_b200.GraspIK(gik_b200.nextage_table(), dev)
n = 1 << 20
g = torch.Generator(device=dev).manual_seed(0)
lo = torch.tensor([0.20, -0.40, 0.93], device=dev, dtype=torch.float64); hi = torch.tensor([0.60, 0.40, 1.40], device=dev, dtype=torch.float64)
pos = lo + torch.rand((n, 3), device=dev, generator=g, dtype=torch.float64) * (hi - lo)
pose = torch.cat([torch.eye(3, device=dev, dtype=torch.float64).reshape(1, 9).expand(n, 9), pos], 1).t().contiguous()
q0 = torch.zeros((15, n), device=dev, dtype=torch.float64)
a = s.solve_soa(q0, pose)
b = s.solve_soa(q0.float(), pose.float())
fa, fb = a[1].bool(), b[1].bool()
both = fa & fb
d = (a[0][:, both] - b[0][:, both].double()).abs().max(dim=0).values
qs = torch.quantile(d[:2000000 if d.numel() > 2000000 else d.numel()].float()[:16000000 // 16], torch.tensor([0.5, 0.9, 0.99, 0.999], device=dev))
print("flag agreement %.6f  (fp64 converged %.4f, fp32 %.4f)" % ((fa == fb).double().mean().item(), fa.double().mean().item(), fb.double().mean().item()))
print("|q32 - q64| over converged: median %.2e  90%% %.2e  99%% %.2e  99.9%% %.2e  max %.2e" % (*qs.tolist(), d.max().item()))
di = (a[2][both] - b[2][both]).abs()
print("iteration count: equal %.4f  within 1: %.4f  max diff %d" % ((di == 0).double().mean().item(), (di <= 1).double().mean().item(), di.max().item()))
print("fp32 residuals of converged: max %.3e (eps 1e-3)" % b[3][:, fb].max().item())
