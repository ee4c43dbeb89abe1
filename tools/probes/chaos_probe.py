"""CPU only: two fp64 restatements of the reference loop (numpy pinv vs C Jacobi-SVD pinv, oracle/) on the first 64
smoke() problems.  Converged problems agree to round-off; problems that FAIL follow a chaotic trajectory and their
returned q differs by O(1) between the two -- the basis of DESIGN.md's statement on failed problems."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import gik_b200
from oracle import c_oracle, grasp_ik_np as gnp

n = 64
rng = np.random.default_rng(0)
P = np.zeros((256, 12)); P[:, [0, 4, 8]] = 1.0
P[:, 9:] = rng.uniform([0.20, -0.40, 0.93], [0.60, 0.40, 1.40], size=(256, 3))
P = P[:n]
q_c, ok_c, it_c, _ = c_oracle.solve(gik_b200.nextage_table().to_c(), np.zeros((n, 15)), P)
rows = []
for i in range(n):
    q, ok = gnp.computeqgrasppose(np.zeros(15), np.eye(3), P[i, 9:])[:2]
    rows.append((i, bool(ok_c[i]), bool(ok), float(np.abs(q - q_c[i]).max())))
conv = [r for r in rows if r[1] and r[2]]
fail = [r for r in rows if not r[1] or not r[2]]
print(f"{len(conv)} problems converged in both: max |q_numpy - q_C| = {max(r[3] for r in conv):.2e}")
print(f"{len(fail)} problems failed in at least one: flags differ on {sum(r[1] != r[2] for r in fail)}, "
      f"|q_numpy - q_C| median {np.median([r[3] for r in fail]):.2e}, max {max(r[3] for r in fail):.2e}, "
      f"{sum(r[3] > 1e-6 for r in fail)} of them differ by more than 1e-6 (indices {[r[0] for r in fail if r[3] > 1e-6]})")
