import sys, time, torch
sys.path.insert(0, '.')
import gik_b200
dev = torch.device("cuda:0")
s = gik_b200.GraspIK(gik_b200.nextage_table(), dev)
n = 1 << 20
g = torch.Generator().manual_seed(0)
pos = torch.tensor([0.2, -0.4, 0.93]) + torch.rand((n, 3), generator=g) * torch.tensor([0.4, 0.8, 0.47])
pose = torch.cat([torch.eye(3).reshape(1, 9).expand(n, 9), pos], 1).contiguous().pin_memory()
q0 = torch.zeros((n, 15)).pin_memory()
for ch in (1, 2, 3, 4, 6, 8, 16):
    for _ in range(2):
        s.solve_host(q0, pose, slabs=ch)
    t = time.perf_counter()
    for _ in range(5):
        s.solve_host(q0, pose, slabs=ch)
    dt = (time.perf_counter() - t) / 5
    print(f"slabs {ch:2d}: {dt*1e3:7.2f} ms  {n/dt/1e6:6.2f} M solves/s")
# device-resident reference
qd = q0.to(dev).t().contiguous(); pd = pose.to(dev).t().contiguous()
for _ in range(2): s.solve_soa(qd, pd)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5): s.solve_soa(qd, pd)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print(f"device-resident: {dt*1e3:7.2f} ms")
