"""Full success predicate on the device (gik_solve_success_*): throughput on 2^20 config-2 problems with and without the
reference's keep-descending-while-colliding tail, and what the tail kernel did (persistent / replayed problems)."""
import sys, time
import torch
sys.path.insert(0, '.')
import gik_b200

dev = torch.device("cuda:0")
s = gik_b200.GraspIK(gik_b200.nextage_table(), dev).attach_scene()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
g = torch.Generator(device=dev).manual_seed(1000)
lo = torch.tensor([0.20, -0.40, 0.93], device=dev); hi = torch.tensor([0.60, 0.40, 1.40], device=dev)
pos = lo + torch.rand((n, 3), device=dev, generator=g) * (hi - lo)
for dtype in (torch.float32, torch.float64):
    pose = torch.cat([torch.eye(3, device=dev).reshape(1, 9).expand(n, 9), pos], 1).t().contiguous().to(dtype)
    q0 = torch.zeros((15, n), device=dev, dtype=dtype)
    for descend in (False, True):
        for _ in range(2):
            out = s.solve_success_soa(q0, pose, descend_while_colliding=descend, return_stats=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = s.solve_success_soa(q0, pose, descend_while_colliding=descend, return_stats=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        q, succ, conv, it, res, st = out
        print(f"{str(dtype):14s} descend={descend!s:5s} {ms:8.2f} ms  {n / ms / 1e3:7.2f} M success-flagged solves/s  success {succ.float().mean():.4f} "
              f"converged-flag {conv.float().mean():.4f}  tail: persistent {int(st[0])} replayed {int(st[1])} replay-iterations {int(st[2])} "
              f"replays-succeeded {int(st[3])}")
    plain = s.solve_soa(q0, pose)
    torch.cuda.synchronize()
    e0.record(); plain = s.solve_soa(q0, pose); e1.record(); torch.cuda.synchronize()
    print(f"{str(dtype):14s} plain solve {e0.elapsed_time(e1):8.2f} ms; converged {plain[1].float().mean():.4f}")
# one drop-in call, worst case: converged but colliding to the cap
import numpy as np
idx = torch.nonzero(conv.bool().logical_not() & (it == 1000) & (plain[1].bool()))[:1]
if idx.numel():
    p = pos[int(idx[0])].cpu().numpy().astype(float)
    for dtype in (torch.float64, torch.float32):
        for _ in range(3):
            r = gik_b200.computeqgrasppose(s, np.zeros(15), None, (np.eye(3), p), dtype=dtype, return_info=True)
        t = time.perf_counter()
        for _ in range(20):
            r = gik_b200.computeqgrasppose(s, np.zeros(15), None, (np.eye(3), p), dtype=dtype, return_info=True)
        print(f"drop-in call on a converged-but-colliding problem ({dtype}): {(time.perf_counter() - t) / 20 * 1e3:.3f} ms wall, success={r[1]}, iterations {r[2]}")
