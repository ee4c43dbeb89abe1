import json, os, sys
import numpy as np, torch
sys.path.insert(0, '.')
import gik_b200
from oracle import c_oracle
solver = gik_b200.GraspIK(gik_b200.nextage_table(), "cuda:0")
rng = np.random.default_rng(0)
n = 256
P = np.zeros((n, 12)); P[:, [0, 4, 8]] = 1.0
P[:, 9:] = rng.uniform([0.20, -0.40, 0.93], [0.60, 0.40, 1.40], size=(n, 3))
gold = json.load(open("tests/golden/ik_golden.json"))
P[0, 9:] = gold["cases"][0]["cube_p"]; P[1, 9:] = gold["cases"][1]["cube_p"]
q_ref, ok_ref, it_ref, r_ref = c_oracle.solve(solver.table.to_c(), np.zeros((n, 15)), P)
for kern in (None, "lane"):
    q64, ok64, info = solver.solve(torch.zeros(15), torch.from_numpy(P), dtype=torch.float64, return_info=True, kernel=kern) if kern else solver.solve(torch.zeros(15), torch.from_numpy(P), dtype=torch.float64, return_info=True)
    ok = ok64.cpu().numpy(); it = info.iters.cpu().numpy(); q = q64.cpu().numpy(); rs = info.resid.cpu().numpy()
    bad = np.nonzero(ok != ok_ref)[0]
    print(os.environ.get("GIK_LIB", "default"), kern, "mismatches", bad, "max|dq| on converged", np.abs(q[ok_ref & ok] - q_ref[ok_ref & ok]).max())
    for i in bad:
        print("  problem", i, "pos", P[i, 9:], "oracle ok/it/resid", ok_ref[i], it_ref[i], r_ref[i], "gpu ok/it/resid", ok[i], it[i], rs[i] if rs.ndim == 2 else rs[:, i])
    d = np.abs(q - q_ref).max(axis=1)
    top = np.argsort(-d)[:5]
    print("  largest |dq|:", [(int(i), float(d[i]), int(it_ref[i]), int(it[i])) for i in top])
