"""Bezier trajectory fit (control.Bezier / control.maketraj, control.py:18-195) -- CPU."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN


@pytest.fixture(scope="module")
def saved():
    return json.load(open(os.path.join(GOLDEN, "trajectory2_reference.json")))


def test_bezier_matches_reference_saved_curves(saved, golden):
    from gik_b200.trajectory import Bezier, load_trajectory_from_json
    q, vq, vvq = load_trajectory_from_json(os.path.join(GOLDEN, "trajectory2_reference.json"))
    # derivative control points as the reference's Bezier.derivative produced and saved them (control.py:49-60, 203-216)
    assert np.abs(np.array(vq.control_points_) - np.array(saved["vq_control_points"])).max() < 1e-12
    assert np.abs(np.array(vvq.control_points_) - np.array(saved["vvq_control_points"])).max() < 1e-12
    T = saved["t_max"]
    # end points are the two recorded grasp configurations; velocity and acceleration vanish at both ends (:110-158)
    assert np.abs(q(0.0) - np.array(golden["cases"][0]["q"])).max() < 1e-15
    assert np.abs(q(T) - np.array(golden["cases"][1]["q"])).max() < 1e-15
    for t in (0.0, T):
        assert np.abs(vq(t)).max() < 1e-12 and np.abs(vvq(t)).max() < 1e-12
    # derivative curves are time derivatives: finite differences of the position curve
    for t in (3.0, 7.7, 12.1):
        hstep = 1e-4
        assert np.abs((q(t + hstep) - q(t - hstep)) / (2 * hstep) - vq(t)).max() < 1e-6
        assert np.abs((vq(t + hstep) - vq(t - hstep)) / (2 * hstep) - vvq(t)).max() < 1e-6
    # Horner evaluation of the reference (control.py:38-47) restated: same values as de Casteljau
    P = np.array(saved["q_control_points"]); n = len(P) - 1
    u = 0.37
    u_op, bc, tn = 1.0 - u, 1, 1
    tmp = P[0] * u_op
    for i in range(1, n):
        tn *= u; bc *= (n - i + 1) / i
        tmp = (tmp + tn * bc * P[i]) * u_op
    assert np.abs((tmp + tn * u * P[-1]) - q(u * T)).max() < 1e-13
    with pytest.raises(ValueError):
        q(T + 1.0)
    with pytest.raises(ValueError):
        Bezier([np.zeros(2)], 0.0, 1.0)


def _path(rng, n_points, dim, q0, q1):
    s = np.linspace(0, 1, n_points)[:, None]
    return q0 + (q1 - q0) * (3 * s ** 2 - 2 * s ** 3) + 0.05 * np.sin(3 * np.pi * s) * rng.normal(size=(1, dim))


def test_maketraj_is_the_constrained_least_squares_optimum():
    from gik_b200.trajectory import bernstein_matrix, maketraj
    rng = np.random.default_rng(0)
    q0, q1 = rng.normal(size=15) * 0.3, rng.normal(size=15) * 0.3
    path = _path(rng, 40, 15, q0, q1)
    (q, vq, vvq), ok = maketraj(q0, q1, path, 15.0)
    assert ok and q.degree_ == 20
    P = np.array(q.control_points_)
    assert np.array_equal(P[0], q0) and np.array_equal(P[1], q0) and np.array_equal(P[2], q0)
    assert np.array_equal(P[-1], q1) and np.array_equal(P[-2], q1) and np.array_equal(P[-3], q1)
    for t in (0.0, 15.0):
        assert np.abs(vq(t)).max() < 1e-12 and np.abs(vvq(t)).max() < 1e-12
    # optimality: the gradient of the reference's cost (control.py:97-105) w.r.t. the free control points vanishes
    B = bernstein_matrix(20, np.linspace(0, 1, 40))
    grad = 2 * B[:, 3:18].T @ (B @ P - path)
    assert np.abs(grad).max() < 1e-9
    cost = ((B @ P - path) ** 2).sum()
    assert cost < 0.15
    # a path the curve cannot follow is reported like the reference does (cost >= 0.15 -> False)
    bad = path + rng.normal(size=path.shape)
    assert maketraj(q0, q1, bad, 15.0)[1] is False


def test_maketraj_agrees_with_the_reference_slsqp_formulation():
    # the reference's own formulation (SLSQP, six equality constraints, control.py:107-170) on a small instance
    from scipy.optimize import minimize
    from gik_b200.trajectory import Bezier, maketraj
    rng = np.random.default_rng(1)
    dim, degree, T = 2, 8, 4.0
    q0, q1 = rng.normal(size=dim), rng.normal(size=dim)
    path = _path(rng, 15, dim, q0, q1)
    times = np.linspace(0, T, len(path))

    def curve(x):
        return Bezier(list(x.reshape(degree + 1, dim)), 0.0, T)

    cost = lambda x: sum(np.sum((curve(x)(t) - qd) ** 2) for t, qd in zip(times, path))
    cons = [{"type": "eq", "fun": lambda x: x.reshape(degree + 1, dim)[0] - q0},
            {"type": "eq", "fun": lambda x: x.reshape(degree + 1, dim)[-1] - q1},
            {"type": "eq", "fun": lambda x: curve(x).derivative(1)(0.0)}, {"type": "eq", "fun": lambda x: curve(x).derivative(1)(T)},
            {"type": "eq", "fun": lambda x: curve(x).derivative(2)(0.0)}, {"type": "eq", "fun": lambda x: curve(x).derivative(2)(T)}]
    res = minimize(cost, np.linspace(q0, q1, degree + 1).flatten(), method="SLSQP", constraints=cons,
                   options={"ftol": 1e-12, "maxiter": 1000})
    assert res.success
    (q, _, _), _ = maketraj(q0, q1, path, T, degree=degree)
    assert np.abs(np.array(q.control_points_) - res.x.reshape(degree + 1, dim)).max() < 1e-4
    assert cost(np.array(q.control_points_).flatten()) <= res.fun + 1e-9


def test_maketraj_batch_equals_single():
    from gik_b200.trajectory import maketraj, maketraj_batch
    rng = np.random.default_rng(2)
    N, n_points, dim = 5, 30, 15
    q0 = rng.normal(size=(N, dim)) * 0.2; q1 = rng.normal(size=(N, dim)) * 0.2
    paths = np.stack([_path(rng, n_points, dim, q0[i], q1[i]) for i in range(N)])
    P, cost = maketraj_batch(torch.from_numpy(q0), torch.from_numpy(q1), torch.from_numpy(paths))
    assert P.shape == (N, 21, dim)
    for i in range(N):
        (q, _, _), ok = maketraj(q0[i], q1[i], paths[i], 15.0)
        ref = np.array(q.control_points_)                                  # the Bernstein design matrix is ill-conditioned
        assert np.abs(P[i].numpy() - ref).max() < 1e-6 * max(1.0, np.abs(ref).max())
        assert ok == bool(cost[i] < 0.15)
