"""Why does bench.py's success-predicate sub-run take longer than tools/probes/tail_probe.py?  Same call, timed with and
without the 256 MB L2 flush write before each step."""
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for mode in ("plain", "flush", "flush+events-inside"):
    for _ in range(2):
        out = s.solve_success_soa(q0, pose, return_stats=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = []
    e0.record()
    for _ in range(3):
        if mode != "plain":
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = s.solve_success_soa(q0, pose, return_stats=True)
        b.record()
        kev.append((a, b))
    e1.record(); torch.cuda.synchronize()
    print(mode, "total/3 %.2f ms" % (e0.elapsed_time(e1) / 3), "inner", ["%.2f" % x.elapsed_time(y) for x, y in kev], out[5].tolist())
