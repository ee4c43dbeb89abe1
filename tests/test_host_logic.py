"""Host-side helpers of the drop-in layer (pose normalisation, SE3 interpolation bookkeeping, sharding) -- CPU."""
import numpy as np
import torch

from conftest import rot_rpy


def test_as_pose12_accepts_every_documented_layout():
    import gik_b200
    R = rot_rpy(0.1, -0.2, 0.3)
    p = np.array([0.3, -0.1, 1.0])
    M = np.eye(4); M[:3, :3] = R; M[:3, 3] = p
    ref = np.concatenate([R.reshape(9), p])
    w = np.sqrt(max(0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    quat = np.array([(R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w), w])
    for x in (ref, M, np.concatenate([p, quat])):
        out = gik_b200.as_pose12(x, dtype=torch.float64, device="cpu")
        assert out.shape == (1, 12) and np.abs(out[0].numpy() - ref).max() < 1e-12
    out = gik_b200.as_pose12(p[None], dtype=torch.float64, device="cpu")
    assert np.abs(out[0].numpy() - np.concatenate([np.eye(3).reshape(9), p])).max() == 0


def test_se3_interpolate_matches_oracle(c_oracle):
    import gik_b200
    rng = np.random.default_rng(0)
    A = np.zeros((16, 12)); B = np.zeros((16, 12))
    for i in range(16):
        A[i, :9] = rot_rpy(*rng.uniform(-1, 1, 3)).reshape(9); A[i, 9:] = rng.uniform(-1, 1, 3)
        B[i, :9] = rot_rpy(*rng.uniform(-1, 1, 3)).reshape(9); B[i, 9:] = rng.uniform(-1, 1, 3)
    B[0] = A[0]                                      # zero displacement
    B[1, :9] = A[1, :9]                              # pure translation (the reference's case, path.py:47)
    for alpha in (0.0, 0.25, 1.0):
        ref = c_oracle.interpolate(A, B, np.full(16, alpha))
        out = gik_b200.se3_interpolate(torch.from_numpy(A), torch.from_numpy(B), alpha).numpy()
        assert np.abs(out - ref).max() < 1e-12
    assert np.abs(c_oracle.interpolate(A, B, np.ones(16)) - B).max() < 1e-12
    lerp = A[1, 9:] + 0.25 * (B[1, 9:] - A[1, 9:])
    assert np.abs(c_oracle.interpolate(A[1:2], B[1:2], np.array([0.25]))[0, 9:] - lerp).max() < 1e-14


def test_edge_num_steps_follows_reference_formula():
    import gik_b200
    a = torch.tensor([[1., 0, 0, 0, 1, 0, 0, 0, 1, 0.33, -0.3, 0.93]], dtype=torch.float64)
    b = torch.tensor([[1., 0, 0, 0, 1, 0, 0, 0, 1, 0.33, -0.3 + 0.1, 0.93]], dtype=torch.float64)
    assert gik_b200.edge_num_steps(a, b).tolist() == [int(0.1 / 0.025) + 1]       # path.py:129-130
    assert gik_b200.edge_num_steps(a, a).tolist() == [1]


def test_shard_bounds_cover_and_balance():
    from gik_b200.dist import shard_bounds
    for n in (0, 1, 7, 64, 1000003):
        for w in (1, 2, 4, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_joint_limit_helpers_need_no_gpu(table):
    # tools.py:11-22 with the reference's names; limits come from the table / model, not from a device handle
    from gik_b200 import tools
    q = np.zeros(15); q[4] = 5.0; q[7] = -9.0
    p = tools.projecttojointlimits(None, q)
    assert p[4] == table.upper[4] and p[7] == table.lower[7] and np.array_equal(p[[0, 1, 2]], [0, 0, 0])
    assert tools.jointlimitsviolated(table, q) and not tools.jointlimitsviolated(table, p)
    assert abs(tools.jointlimitscost(table, q) - max(5.0 - table.upper[4], table.lower[7] + 9.0)) < 1e-12


def test_tools_dropin_signatures_cpu():
    # tools.py helpers with the reference's call shapes, on a duck-typed stand-in for the pinocchio cube wrapper
    # (tools.py:54-59: getcubeplacement takes the CUBE) -- no GPU, no pinocchio
    import inspect
    from gik_b200 import tools

    class SE3:
        def __init__(self, M): self.M = np.asarray(M, float)
        def __mul__(self, o): return SE3(self.M @ o.M)

    class Model:
        def existFrame(self, name): return name in ("LARM_HOOK", "RARM_HOOK")
        def getFrameId(self, name): return {"LARM_HOOK": 1, "RARM_HOOK": 2}[name]

    class Geom: pass
    class Bag: pass
    place = np.eye(4); place[:3, 3] = [0.33, -0.3, 0.93]
    hookL = np.eye(4); hookL[:3, 3] = [0.0, 0.05, 0.0]
    cube = Bag(); cube.model = Model(); cube.data = Bag(); cube.data.oMf = [SE3(np.eye(4)), SE3(hookL), SE3(np.eye(4))]
    g = Geom(); g.placement = SE3(place)
    cube.collision_model = Bag(); cube.collision_model.geometryObjects = [g]
    assert tools._is_cube_wrapper(cube) and not tools._is_cube_wrapper(None) and not tools._is_cube_wrapper(object())
    assert np.array_equal(tools.getcubeplacement(cube).M, place)
    assert np.allclose(tools.getcubeplacement(cube, "LARM_HOOK").M[:3, 3], [0.33, -0.25, 0.93])
    assert np.array_equal(g.placement.M, place)                    # the stored placement is not modified
    # same parameter names / order as the reference for the positional part of every helper
    for name, lead in (("jointlimitscost", ["robot", "q"]), ("jointlimitsviolated", ["robot", "q"]),
                       ("projecttojointlimits", ["robot", "q"]), ("collision", ["robot", "q"]),
                       ("distanceToObstacle", ["robot", "q"]), ("setcubeplacement", ["robot", "cube", "oMf"])):
        assert list(inspect.signature(getattr(tools, name)).parameters)[:len(lead)] == lead
    assert list(inspect.signature(tools.getcubeplacement).parameters)[1] == "hookname"
    # joint-limit helpers work without a device handle
    t = gik_b200_table()
    q = np.zeros(15); q[3] = 9.0
    assert tools.jointlimitsviolated(t, q) and not tools.jointlimitsviolated(t, np.zeros(15))
    assert tools.projecttojointlimits(t, q)[3] == t.upper[3]


def gik_b200_table():
    import gik_b200
    return gik_b200.nextage_table()


# ----------------------------------------------------------------------------------------------------------
# The fused all-gather's bookkeeping (csrc/gik_kernels.cu: queue_take, mark_update / mark_publish, the lookout's F),
# restated in Python and driven through random interleavings: the device code is these few integer rules plus
# atomics, and the properties below are what makes pushing a chunk before the kernel ends safe.
# ----------------------------------------------------------------------------------------------------------
def _queue_take(state, n, k, from_top):
    """One warp refill on the two-ended queue word (low half: handed out from the bottom, high half: from the top).
    Returns the work items of the k requesting lanes (-1 = none) and whether the warp now knows the queue is exhausted."""
    b, t = state
    if from_top:
        state[1] = t + k
        cands = [n - 1 - t - r for r in range(k)]
        items = [c if c >= b else -1 for c in cands]
    else:
        state[0] = b + k
        cands = [b + r for r in range(k)]
        items = [c if c < n - t else -1 for c in cands]
    return items, (b + t + k >= n)


def test_two_ended_queue_hands_every_problem_out_exactly_once():
    rng = np.random.default_rng(0)
    for trial in range(300):
        n = int(rng.integers(1, 400))
        state = [0, 0]
        seen = []
        exhausted = 0
        for _ in range(10 * n + 50):                       # refills keep coming after exhaustion, like late warps
            k = int(rng.integers(1, 33))
            items, ex = _queue_take(state, n, k, bool(rng.integers(0, 2)))
            seen += [i for i in items if i >= 0]
            exhausted += ex
            if len(seen) == n and rng.random() < 0.3:
                break
        assert sorted(seen) == list(range(n)), (trial, n)
        assert exhausted >= 1                              # somebody learned that the queue ran dry


def test_low_water_marks_never_overstate_the_finished_prefix():
    # W warps of L lanes; bottom warps publish their low-water mark ONE finish event late, top warps publish "done" at
    # once; the lookout's F = min(marks << shift, bottom count, n - top count) must never exceed the true finished prefix
    # at any moment, under random finish orders and random delays between computing and publishing a mark.
    rng = np.random.default_rng(1)
    DONE, shift = 1 << 30, 2
    for trial in range(60):
        n, W, L = int(rng.integers(50, 600)), int(rng.integers(2, 7)), 4
        top = [bool(w >= W // 2 and rng.random() < 0.8) for w in range(W)]
        state = [0, 0]
        finished = np.zeros(n, bool)
        lanes = [[-1] * L for _ in range(W)]               # current problem of every lane (-1: none)
        marks = [DONE if top[w] else 0 for w in range(W)]  # what the lookout can read
        pending = [0] * W
        exhausted = [False] * W
        gone = [False] * W

        def check():
            m = min(marks)
            F = n if m == DONE else (m << shift)
            F = min(F, state[0], n - state[1], n)
            prefix = int(np.argmin(finished)) if not finished.all() else n
            assert F <= prefix, (trial, F, prefix)

        def refill(w):
            need = [j for j in range(L) if lanes[w][j] < 0]
            if need and not exhausted[w]:
                items, ex = _queue_take(state, n, len(need), top[w])
                exhausted[w] = ex
                for j, it in zip(need, items):
                    lanes[w][j] = it

        for w in range(W):
            refill(w)
        steps = 0
        while not all(gone):
            steps += 1
            assert steps < 100000
            w = int(rng.integers(0, W))
            if gone[w]:
                continue
            active = [j for j in range(L) if lanes[w][j] >= 0]
            if not active:
                if exhausted[w]:
                    marks[w] = DONE; gone[w] = True        # mark_exit (after a fence: everything it stored is visible)
                else:
                    refill(w)
                check()
                continue
            # a finish event of this warp: publish the PREVIOUS mark, store results, retire, compute the next mark, refill
            if not top[w] and pending[w] != marks[w]:
                marks[w] = pending[w]
            check()
            for j in active:
                if rng.random() < 0.5 or len(active) == 1:
                    finished[lanes[w][j]] = True
                    lanes[w][j] = -1
            still = [lanes[w][j] >> shift for j in range(L) if lanes[w][j] >= 0]
            if still:
                pending[w] = min(still)
            check()
            refill(w)
            check()
        assert finished.all()
        check()
