"""The device arithmetic (csrc/gik_core.cuh) compiled for the host, against the oracle -- CPU only.  This checks the
kernel MATH (backward chain walk, block-structured Cholesky + Sherman-Morrison step, log6) in fp64 and fp32
without a GPU; the `-m gpu` tests check the kernels themselves through the C ABI."""
import numpy as np

from conftest import make_poses, rot_rpy


def test_hostsim_goldens(golden, table_c, hostsim):
    P = np.array([c["cube_R"] + c["cube_p"] for c in golden["cases"]], float)
    for dt, tol in ((np.float64, 1e-9), (np.float32, 1e-4)):
        q, ok, it, r = hostsim.solve(table_c, np.zeros((2, 15)), P, dt)
        assert ok.all() and (r < 1e-3).all()
        for i, c in enumerate(golden["cases"]):
            assert np.abs(q[i] - np.array(c["q"])).max() < tol
            assert abs(int(it[i]) - c["iterations_chart"]) <= (0 if dt == np.float64 else 1)


def test_hostsim_fk_jac_tolerances(table, table_c, hostsim, c_oracle):
    rng = np.random.default_rng(0)
    Q = rng.uniform(table.lower, table.upper, size=(256, 15))
    R, p = c_oracle.fk(table_c, Q)
    J = c_oracle.jac(table_c, Q)
    for dt, tol in ((np.float64, 1e-9), (np.float32, 1e-5)):   # north_star tolerances
        R2, p2 = hostsim.fk(table_c, Q, dt)
        J2 = hostsim.jac(table_c, Q, dt)
        assert np.abs(R - R2).max() < tol and np.abs(p - p2).max() < tol and np.abs(J - J2).max() < tol


def test_hostsim_solve_matches_oracle(table_c, hostsim, c_oracle):
    P = make_poses(96, 11)
    P[:8, :9] = rot_rpy(0, 0, 0.4).reshape(9)          # a few yawed cubes
    P[:, 9] = np.minimum(P[:, 9], 0.55)                # stay clear of the near-singular rim (see test_gpu_parity._assert_q_close)
    qo, oko, ito, _ = c_oracle.solve(table_c, np.zeros((96, 15)), P)
    q, ok, it, r = hostsim.solve(table_c, np.zeros((96, 15)), P, np.float64)
    assert (ok == oko).all()
    assert np.abs(q[oko] - qo[oko]).max() < 1e-9 and (it[oko] == ito[oko]).all()
    q, ok, it, r = hostsim.solve(table_c, np.zeros((96, 15)), P, np.float32)
    assert (ok == oko).mean() >= 0.97
    both = ok & oko
    assert np.abs(q[both] - qo[both]).max() < 1e-3 and np.abs(it[both] - ito[both]).max() <= 2
    assert (r[ok] < 1e-3).all()


def test_hostsim_damping_shortens_step(table_c, hostsim):
    P = make_poses(4, 2)
    q0, *_ = hostsim.solve(table_c, np.zeros((4, 15)), P, np.float64, max_iters=1)
    q1, *_ = hostsim.solve(table_c, np.zeros((4, 15)), P, np.float64, max_iters=1, damping=1.0)
    assert (np.abs(q1).sum(1) < np.abs(q0).sum(1)).all()


def test_hostsim_translation_sparsity_specialisation_is_exact(table_c, hostsim):
    # the Nextage-pattern instantiation only drops FMAs whose translation operand is exactly 0: same bits
    P = make_poses(16, 9)
    for dt in (np.float64, np.float32):
        a = hostsim.solve(table_c, np.zeros((16, 15)), P, dt, max_iters=50)
        b = hostsim.solve(table_c, np.zeros((16, 15)), P, dt, max_iters=50, flags=1)     # generic instantiation
        assert np.abs(a[0] - b[0]).max() < (1e-13 if dt == np.float64 else 1e-5)


def test_hostsim_random_tables_of_the_supported_topology(table, hostsim, c_oracle):
    # kernel math vs oracle on robots of the same topology but random dimensions / frames / limits (generic TZ = 0 path)
    import copy
    from conftest import rot_rpy
    rng = np.random.default_rng(5)
    for trial in range(4):
        t = copy.deepcopy(table)
        t.joint_p = t.joint_p * rng.uniform(0.85, 1.15, size=t.joint_p.shape) + rng.normal(scale=0.01, size=t.joint_p.shape)
        t.hand_R = np.array([rot_rpy(*rng.uniform(-0.3, 0.3, 2), 1.5 + rng.normal(scale=0.3)) for _ in range(2)])
        t.hand_p = t.hand_p + rng.normal(scale=0.01, size=(2, 3))
        t.hook_R = np.array([rot_rpy(0, 0, rng.normal(scale=0.1)), rot_rpy(0, 0, -3.1 + rng.normal(scale=0.1))])
        tc = t.to_c()
        Q = rng.uniform(t.lower, t.upper, size=(32, 15))
        R, p = c_oracle.fk(tc, Q)
        J = c_oracle.jac(tc, Q)
        R2, p2 = hostsim.fk(tc, Q, np.float64)
        J2 = hostsim.jac(tc, Q, np.float64)
        assert np.abs(R - R2).max() < 1e-12 and np.abs(p - p2).max() < 1e-12 and np.abs(J - J2).max() < 1e-12
        P = make_poses(24, 100 + trial)
        P[:, 9] = np.minimum(P[:, 9], 0.5)
        qo, oko, ito, _ = c_oracle.solve(tc, np.zeros((24, 15)), P)
        q, ok, it, r = hostsim.solve(tc, np.zeros((24, 15)), P, np.float64)
        assert (ok == oko).all()
        both = ok & oko
        if both.any():
            assert np.quantile(np.abs(q[both] - qo[both]).max(axis=1), 0.9) < 1e-9 and (it[both] == ito[both]).mean() > 0.9


def test_hostsim_damped_restarts_match_oracle(table, table_c, hostsim, c_oracle):
    # config 3's setting: damping 1e-6, initial q uniform in [lower, upper] over all 15 joints (restart 0 = zeros)
    rng = np.random.default_rng(13)
    n = 64
    P = np.repeat(make_poses(8, 14), 8, axis=0)
    Q0 = rng.uniform(table.lower, table.upper, size=(n, 15)); Q0[::8] = 0.0
    qo, oko, ito, ro = c_oracle.solve(table_c, Q0, P, damping=1e-6)
    q, ok, it, r = hostsim.solve(table_c, Q0, P, np.float64, damping=1e-6)
    assert (ok == oko).mean() >= 0.95
    both = ok & oko
    assert both.sum() >= 8
    d = np.abs(q[both] - qo[both]).max(axis=1)
    assert np.quantile(d, 0.9) < 1e-8 and (it[both] == ito[both]).mean() >= 0.9
    assert np.abs(r[both] - ro[both]).max() < 1e-9


def test_hostsim_nextage_table_takes_the_specialised_wrist_step(table, table_c, hostsim):
    # the launcher's fast path needs every bit of the Nextage pattern: zero translations, spherical wrist, and hand frames
    # that are rotations about the tip joint's axis (kTipZ, bit 24); a hand frame tilted off that axis must lose the bit
    # (and with it the specialised step -- the generic one then runs, same results: next test)
    import copy, ctypes
    f = hostsim.lib.hostsim_table_pattern
    f.restype = ctypes.c_longlong
    nextage = (3 << 0) | (3 << 6) | (1 << 9) | (1 << 13) | (6 << 15) | (3 << 18) | (1 << 24)
    pat = f(ctypes.byref(table_c))
    assert pat >= 0 and (pat & nextage) == nextage
    t2 = copy.deepcopy(table)
    t2.hand_R = np.array(t2.hand_R, float).copy()
    t2.hand_R[0] = rot_rpy(0.1, 0.0, 1.5708)
    pat2 = f(ctypes.byref(t2.to_c()))
    assert pat2 >= 0 and not (pat2 & (1 << 24)) and (pat2 & (nextage & ~(1 << 24))) == (nextage & ~(1 << 24))


def test_hostsim_tilted_hand_frame_uses_the_generic_step(table, hostsim, c_oracle):
    # same robot with the left hand frame tilted off the tip axis: the builder drops kTipZ, the generic path solves it,
    # and the result still matches the oracle
    import copy
    t2 = copy.deepcopy(table)
    t2.hand_R = np.array(t2.hand_R, float).copy()
    t2.hand_R[0] = rot_rpy(0.1, 0.0, 1.5708)
    tc2 = t2.to_c()
    P = make_poses(24, 13)
    P[:, 9] = np.minimum(P[:, 9], 0.5)
    qo, oko, ito, _ = c_oracle.solve(tc2, np.zeros((24, 15)), P)
    q, ok, it, r = hostsim.solve(tc2, np.zeros((24, 15)), P, np.float64)
    assert (ok == oko).all() and oko.any()
    assert np.abs(q[oko] - qo[oko]).max() < 1e-8 and (it[oko] == ito[oko]).all()


def test_hostsim_final_kernel_math_against_the_oracle_2048_problems(table_c, hostsim, c_oracle):
    # the arithmetic of the final kernels (tip-aligned spherical-wrist step; packed fp32 log6 without the small-angle series)
    # on 2,048 workspace problems against the oracle, on the CPU: flags, iteration counts, configurations
    n = 2048
    P = make_poses(n, 2025)
    qo, oko, ito, _ = c_oracle.solve(table_c, np.zeros((n, 15)), P)
    q, ok, it, r = hostsim.solve(table_c, np.zeros((n, 15)), P, np.float64)
    assert (ok != oko).sum() <= 2                                   # (chaotic failing trajectories: DESIGN.md "Sensitivity")
    both = ok & oko
    d = np.abs(q[both] - qo[both]).max(axis=1)
    assert np.quantile(d, 0.99) < 1e-9 and (it[both] == ito[both]).mean() >= 0.995 and (r[ok] < 1e-3).all()
    q, ok, it, r = hostsim.solve(table_c, np.zeros((n, 15)), P, np.float32)
    assert (ok != oko).sum() <= 4                                   # >= 99.8 %
    both = ok & oko
    d = np.abs(q[both] - qo[both]).max(axis=1)
    assert np.quantile(d, 0.995) < 1e-3 and np.quantile(np.abs(it[both] - ito[both]), 0.995) <= 2 and (r[ok] < 1e-3).all()
    assert 0.4 < oko.mean() < 0.9
