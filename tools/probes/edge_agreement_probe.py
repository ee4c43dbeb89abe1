"""How often do the fp32 and fp64 edge marches stop at different steps, and on which edges?  (The fp64 march equals the C
oracle step for step: tests/test_gpu_parity.py::test_project_edges_matches_oracle.)  Also lane vs pair mapping in fp32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import gik_b200
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from conftest import make_poses, rot_rpy

solver = gik_b200.GraspIK(gik_b200.nextage_table(), "cuda:0")
def t(a, dt): return torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device="cuda:0")
for E, S, spread in ((4096, 6, 0.08), (4096, 16, 0.25)):
    A = make_poses(E, 51, "sampler"); A[:, 11] = 0.93 + 0.2 * np.random.default_rng(1).uniform(size=E)
    B = A.copy(); B[:, 9:] += np.random.default_rng(2).uniform(-spread, spread, size=(E, 3))
    B[::4, :9] = rot_rpy(0, 0, 0.5).reshape(9)
    q0, ok0 = solver.solve(torch.zeros(15), t(A, torch.float64), dtype=torch.float64)
    keep = ok0.cpu().numpy()
    A, B, q0 = A[keep], B[keep], q0[keep]
    n = len(A)
    ns = np.full(n, S, np.int32)
    res = {}
    for name, dt, kern in (("f64", torch.float64, None), ("f32", torch.float32, None), ("f32-lane", torch.float32, "lane")):
        nsd = torch.as_tensor(ns, device="cuda:0")
        p, nv, it = solver.project_edges_soa(q0.to(dt).t().contiguous(), t(A, dt).t().contiguous(), t(B, dt).t().contiguous(), nsd, S, kernel=kern)
        res[name] = (p.double().cpu().numpy(), nv.cpu().numpy(), it.cpu().numpy())
    for a, b in (("f32", "f64"), ("f32-lane", "f32")):
        nva, nvb = res[a][1], res[b][1]
        same = nva == nvb
        print(f"E={n} S={S} spread={spread}: {a} vs {b}: n_valid equal on {same.mean():.5f} ({(~same).sum()} differ); fully projected {np.mean(nvb == S):.3f}")
        # the differing edges: iterations spent on the step where the shorter march stopped (a failing step costs 1000)
        for e in np.nonzero(~same)[0][:12]:
            print(f"   edge {e}: n_valid {nva[e]} vs {nvb[e]}, total iterations {res[a][2][e]} vs {res[b][2][e]}")
        k = np.minimum(nva, nvb)
        d = max(np.abs(res[a][0][:k[e], :, e] - res[b][0][:k[e], :, e]).max(initial=0) for e in range(n))
        print(f"   max |q| difference over the steps both marched: {d:.3e}")
