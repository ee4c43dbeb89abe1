import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np, torch
import gik_b200
from conftest import make_poses
from oracle import c_oracle
table = gik_b200.nextage_table(); tc = table.to_c()
solver = gik_b200.GraspIK(table, "cuda:0")
for n, scale in ((64, 0.3), (1024, 0.3), (1024, 1.0)):
    rng = np.random.default_rng(5)
    P = make_poses(n, 31)
    Q0 = rng.uniform(table.lower, table.upper, size=(n, 15)) * scale
    Q0[:, 1] = 2.0
    qo, oko, ito, _ = c_oracle.solve(tc, Q0, P)
    q, ok, info = solver.solve(torch.as_tensor(Q0, device="cuda:0"), torch.as_tensor(P, device="cuda:0"), dtype=torch.float64, return_info=True)
    q = q.cpu().numpy(); ok = ok.cpu().numpy(); it = info.iters.cpu().numpy()
    both = ok & oko
    d = np.abs(q[both] - qo[both]).max(axis=1)
    print(n, scale, "flags differ", (ok != oko).sum(), "converged", oko.mean(), "d quantiles 50/90/99/max", np.quantile(d, [0.5, 0.9, 0.99, 1.0]), "iters equal", (it[both] == ito[both]).mean())
