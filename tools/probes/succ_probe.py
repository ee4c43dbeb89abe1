import sys, time, torch
sys.path.insert(0, '.')
import gik_b200
dev = torch.device("cuda:0")
s = gik_b200.GraspIK(gik_b200.nextage_table(), dev).attach_scene()
n = 1 << 20
g = torch.Generator(device=dev).manual_seed(0)
lo = torch.tensor([0.20, -0.40, 0.93], device=dev); hi = torch.tensor([0.60, 0.40, 1.40], device=dev)
pos = lo + torch.rand((n, 3), device=dev, generator=g) * (hi - lo)
pose = torch.cat([torch.eye(3, device=dev).reshape(1, 9).expand(n, 9), pos], 1).t().contiguous()
q0 = torch.zeros((15, n), device=dev)
def T(f, reps=3):
    f(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): r = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / reps * 1e3, r
t, (q, conv, it, res) = T(lambda: s.solve_soa(q0, pose)); print("solve", t)
convb = conv.bool()
t, idx = T(lambda: torch.nonzero(convb).flatten()); print("nonzero", t, idx.numel())
t, qs = T(lambda: q[:, idx].contiguous()); print("gather q", t)
t, ps = T(lambda: pose[:, idx].contiguous()); print("gather pose", t)
t, c = T(lambda: s.collision_soa(qs, ps)); print("collision subset", t)
t, c2 = T(lambda: s.collision_soa(q, pose)); print("collision all", t)
t, _ = T(lambda: s.solve_success_soa(q0, pose, descend_while_colliding=False)); print("solve_success", t)
