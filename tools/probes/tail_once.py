"""One gik_solve_success_f32 call on 2^20 config-2 problems (for an ncu launch list: which kernel of the call costs what)."""
import sys
import torch
sys.path.insert(0, '.')
import gik_b200
dev = torch.device("cuda:0")
s = gik_b200.GraspIK(gik_b200.nextage_table(), dev).attach_scene()
n = 1 << 20
g = torch.Generator(device=dev).manual_seed(1000)
lo = torch.tensor([0.20, -0.40, 0.93], device=dev); hi = torch.tensor([0.60, 0.40, 1.40], device=dev)
pos = lo + torch.rand((n, 3), device=dev, generator=g) * (hi - lo)
pose = torch.cat([torch.eye(3, device=dev).reshape(1, 9).expand(n, 9), pos], 1).t().contiguous()
q0 = torch.zeros((15, n), device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(reps):
    out = s.solve_success_soa(q0, pose, return_stats=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = s.solve_success_soa(q0, pose, return_stats=True); e1.record(); torch.cuda.synchronize()
print("one call:", e0.elapsed_time(e1), "ms; stats", out[5].tolist())
