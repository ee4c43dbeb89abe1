"""CPU ORACLE (numpy, fp64) -- TEST INFRASTRUCTURE ONLY, never on the product path.

A dependency-free restatement of the reference's dual-arm grasp IK
(`/root/reference/inverse_geometry.py:17-100`) and of the pinocchio primitives it
calls.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module, and only as the checker.

Parity status: PINNED.  The restatement reproduces the reference's own golden
outputs (`trajectory.json` q_control_points[0] / [-1], fixed there by
control.py:110-119,435-437) to ~1e-14 in 740 / 736 iterations, and the notebook
known answer for `oMf[LARM_EFF]` at q0 (lab_instructions.ipynb:290-293); see
`tests/test_oracle.py` and `tests/golden/`.

The arithmetic itself lives in un-vendored third-party code: pinocchio (PyPI `pin`,
version not pinned by requirements.txt:1; notebooks show a 2.x build) and numpy's
LAPACK `pinv`.  What is restated here is pinocchio's published algorithm:
  * framesForwardKinematics: oMi[i] = oMi[parent] * jointPlacement[i] * Rot(axis_i, q_i),
    oMf[f] = oMi[parent(f)] * placement(f)            (inverse_geometry.py:58)
  * computeFrameJacobian, default LOCAL reference frame  (inverse_geometry.py:75-76)
  * log6 / log3                                          (inverse_geometry.py:66-67)
  * integrate on a vector-space configuration            (inverse_geometry.py:86)
Every function cites the reference line it follows.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# Model constants, typed independently of the product's flattener from
# models/nextagea_description/urdf/NextageaOpen.urdf:580-730, models/cubes/cube_small.urdf:34-48,
# config.py:22-37 and setup_pinocchio.py:28-32.  q index = pinocchio joint id - 1; joint order
# as printed in lab_instructions.ipynb:210-226.
# --------------------------------------------------------------------------------------
NQ = 15
JOINT_NAMES = [
    "CHEST_JOINT0", "HEAD_JOINT0", "HEAD_JOINT1",
    "LARM_JOINT0", "LARM_JOINT1", "LARM_JOINT2", "LARM_JOINT3", "LARM_JOINT4", "LARM_JOINT5",
    "RARM_JOINT0", "RARM_JOINT1", "RARM_JOINT2", "RARM_JOINT3", "RARM_JOINT4", "RARM_JOINT5",
]
# parent q index (-1 = universe)
PARENT = np.array([-1, 0, 1, 0, 3, 4, 5, 6, 7, 0, 9, 10, 11, 12, 13], dtype=np.int64)
# 0 = x, 1 = y, 2 = z
AXIS = np.array([2, 2, 1, 2, 1, 1, 0, 1, 2, 2, 1, 1, 0, 1, 2], dtype=np.int64)
ROBOT_Z = 0.85  # config.py:33 ROBOT_PLACEMENT, applied to jointPlacements[1] (setup_pinocchio.py:32)
TRANS = np.array([
    [0.0, 0.0, 0.267 + ROBOT_Z],
    [0.0, 0.0, 0.302],
    [0.0, 0.0, 0.08],
    [0.04, 0.135, 0.1015],
    [0.0, 0.0, 0.066],
    [0.0, 0.095, -0.25],
    [0.1805, 0.0, -0.03],
    [0.1495, 0.0, 0.0],
    [0.0, 0.0, -0.1335],
    [0.04, -0.135, 0.1015],
    [0.0, 0.0, 0.066],
    [0.0, -0.095, -0.25],
    [0.1805, 0.0, -0.03],
    [0.1495, 0.0, 0.0],
    [0.0, 0.0, -0.1335],
])
LOWER = np.array([-3.14159, -1.22173, -0.401425,
                  -1.5707963, -2.44346, -1.22173, -3.1415926, -3.57792, -2.7123889,
                  -1.570796, -2.44346, -1.22173, -1.74532, -3.5779, -2.712388])
UPPER = np.array([3.14159, 1.22173, 1.308997,
                  1.5707963, 1.0471975, 1.5707963, 1.7453292, 1.134464, 2.7123889,
                  1.570796, 1.047197, 1.570796, 3.141592, 1.134464, 2.712388])


def rotz(a: float) -> np.ndarray:
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def rot_axis(axis: int, a: float) -> np.ndarray:
    """Elementary rotation about coordinate axis 0/1/2 (pinocchio JointModelRX/RY/RZ)."""
    c, s = np.cos(a), np.sin(a)
    if axis == 0:
        return np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])
    if axis == 1:
        return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


# Effector frames LARM_EFF / RARM_EFF (config.py:25-26): fixed joints on LARM_JOINT5 (q8) and
# RARM_JOINT5 (q14), NextageaOpen.urdf:717-730.  rpy literal 1.5708 is used verbatim.
HAND_JOINT = (8, 14)
HAND_R = (rotz(1.5708), rotz(1.5708))
HAND_P = (np.array([0.082, 0.05, -0.02]), np.array([0.082, -0.05, -0.02]))
# Hook frames LARM_HOOK / RARM_HOOK on the cube (config.py:28-29, cube_small.urdf:34-48); -3.14 verbatim.
HOOK_R = (np.eye(3), rotz(-3.14))
HOOK_P = (np.array([0.0, 0.05, 0.0]), np.array([0.0, -0.05, 0.0]))

EPSILON = 1e-3      # config.py:22
DT = 1e-2           # inverse_geometry.py:54
MAX_ITERS = 1000    # inverse_geometry.py:53
CUBE_PLACEMENT = (np.eye(3), np.array([0.33, -0.3, 0.93]))          # config.py:36
CUBE_PLACEMENT_TARGET = (np.eye(3), np.array([0.4, 0.11, 0.93]))    # config.py:37


# --------------------------------------------------------------------------------------
# pinocchio primitives
# --------------------------------------------------------------------------------------
def forward_kinematics(q):
    """pin.framesForwardKinematics (inverse_geometry.py:58): world placements of the 15 joints."""
    R = np.zeros((NQ, 3, 3))
    p = np.zeros((NQ, 3))
    for i in range(NQ):
        par = PARENT[i]
        Rp, pp = (np.eye(3), np.zeros(3)) if par < 0 else (R[par], p[par])
        p[i] = pp + Rp @ TRANS[i]
        R[i] = Rp @ rot_axis(int(AXIS[i]), q[i])
    return R, p


def hand_placement(q, hand, fk=None):
    """data.oMf[getFrameId(LEFT_HAND / RIGHT_HAND)] (inverse_geometry.py:62-63)."""
    R, p = forward_kinematics(q) if fk is None else fk
    j = HAND_JOINT[hand]
    return R[j] @ HAND_R[hand], p[j] + R[j] @ HAND_P[hand]


def frame_jacobian_local(q, hand, fk=None):
    """pin.computeFrameJacobian(model, data, q, frame_id) with the default LOCAL reference
    frame (inverse_geometry.py:75-76): 6x15, rows [linear; angular], column k non-zero only
    for joints on the frame's support chain."""
    R, p = forward_kinematics(q) if fk is None else fk
    Rf, pf = hand_placement(q, hand, (R, p))
    J = np.zeros((6, NQ))
    k = HAND_JOINT[hand]
    while k >= 0:
        a = R[k][:, AXIS[k]]
        J[:3, k] = Rf.T @ np.cross(a, pf - p[k])
        J[3:, k] = Rf.T @ a
        k = PARENT[k]
    return J


_TS_PREC3 = np.finfo(np.float64).eps ** 0.25   # pinocchio TaylorSeriesExpansion<double>::precision<3>()


def log3(R):
    """pinocchio log3 (used by pin.log at inverse_geometry.py:66-67): returns (omega, theta)."""
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr >= 3.0:
        tr, theta = 3.0, 0.0
    elif tr <= -1.0:
        tr, theta = -1.0, np.pi
    else:
        theta = np.arccos((tr - 1.0) / 2.0)
    if theta >= np.pi - 1e-2:
        cphi = -(tr - 1.0) / 2.0
        beta = theta * theta / (1.0 + cphi)
        tmp = (np.diag(R) + cphi) * beta
        w = np.array([
            (1.0 if R[2, 1] > R[1, 2] else -1.0) * (np.sqrt(tmp[0]) if tmp[0] > 0 else 0.0),
            (1.0 if R[0, 2] > R[2, 0] else -1.0) * (np.sqrt(tmp[1]) if tmp[1] > 0 else 0.0),
            (1.0 if R[1, 0] > R[0, 1] else -1.0) * (np.sqrt(tmp[2]) if tmp[2] > 0 else 0.0),
        ])
    else:
        t = (theta / np.sin(theta) if theta > _TS_PREC3 else 1.0) / 2.0
        w = t * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return w, theta


def log6(R, p):
    """pinocchio log6 -> Motion.vector = [v; w] (inverse_geometry.py:66-67)."""
    w, t = log3(R)
    t2 = t * t
    if t < _TS_PREC3:
        alpha = 1.0 - t2 / 12.0 - t2 * t2 / 720.0
        beta = 1.0 / 12.0 + t2 / 720.0
    else:
        st, ct = np.sin(t), np.cos(t)
        alpha = t * st / (2.0 * (1.0 - ct))
        beta = 1.0 / t2 - st / (2.0 * t * (1.0 - ct))
    v = alpha * p - 0.5 * np.cross(w, p) + (beta * np.dot(w, p)) * w
    return np.concatenate([v, w])


def hook_targets(cube_R, cube_p):
    """tools.getcubeplacement(cube, LEFT_HOOK / RIGHT_HOOK) after setcubeplacement
    (inverse_geometry.py:42-46, tools.py:54-68): cubetarget * hook offset."""
    return [(cube_R @ HOOK_R[h], cube_p + cube_R @ HOOK_P[h]) for h in (0, 1)]


def hand_error(q, hand, target, fk=None):
    """pin.log(oMhand.inverse() * oMcube).vector (inverse_geometry.py:66-67)."""
    Rh, ph = hand_placement(q, hand, fk)
    Rt, pt = target
    return log6(Rh.T @ Rt, Rh.T @ (pt - ph))


def project_to_joint_limits(q):
    """tools.projecttojointlimits (tools.py:21-22)."""
    return np.minimum(np.maximum(LOWER, q), UPPER)


# --------------------------------------------------------------------------------------
# The hot loop
# --------------------------------------------------------------------------------------
def computeqgrasppose(qcurrent, cube_R, cube_p, *, eps=EPSILON, dt=DT, max_iters=MAX_ITERS,
                      collision=None, return_info=False, damping=0.0):
    """inverse_geometry.computeqgrasppose (inverse_geometry.py:17-100), loop order preserved:
    residual check BEFORE the update, post-update q returned on exhaustion, `collision(q)`
    evaluated only when both norms pass (short-circuit at :70) and once more at :97.
    `collision` is a callable q -> bool or None (treated as never colliding).
    `damping` = 0 is the reference (pinv, :83); > 0 replaces the step by the damped least-squares form
    J^T (J J^T + damping I)^-1 e that BASELINE config 3 asks for on random restarts (not a reference setting)."""
    targets = hook_targets(np.asarray(cube_R, float), np.asarray(cube_p, float))
    q = np.array(qcurrent, dtype=np.float64).copy()
    success = False
    iters = max_iters
    nL = nR = np.nan
    for it in range(max_iters):
        fk = forward_kinematics(q)
        eL = hand_error(q, 0, targets[0], fk)
        eR = hand_error(q, 1, targets[1], fk)
        nL, nR = np.linalg.norm(eL), np.linalg.norm(eR)
        if nL < eps and nR < eps and not (collision is not None and collision(q)):
            success = True
            iters = it
            break
        J = np.vstack([frame_jacobian_local(q, 0, fk), frame_jacobian_local(q, 1, fk)])
        if damping > 0.0:
            vq = J.T @ np.linalg.solve(J @ J.T + damping * np.eye(12), np.hstack([eL, eR]))
        else:
            vq = np.linalg.pinv(J) @ np.hstack([eL, eR])      # inverse_geometry.py:83
        q = q + vq * dt                                       # pin.integrate, all-revolute model (:86)
        q = project_to_joint_limits(q)                        # :89
    if collision is not None and collision(q):                # :97-98
        success = False
    if return_info:
        return q, success, iters, np.array([nL, nR])
    return q, success


def project_path(q_curr, cube_a, cube_b, *, step_size=0.025, num_steps=None, **kw):
    """path.project_path (path.py:125-163) for translation-only cube motion (identity rotation,
    as sampled at path.py:47): SE3.Interpolate reduces to a lerp of the translation.  Cube-vs-scene
    collision tests (path.py:144-149) are harness-side and not restated.  Returns the list of
    configurations after q_curr that the reference would append, in order."""
    Ra, pa = cube_a
    Rb, pb = cube_b
    if num_steps is None:
        num_steps = int(np.linalg.norm(pa - pb) / step_size) + 1      # path.py:129-130
    path = []
    q = np.asarray(q_curr, float)
    for step in range(1, num_steps + 1):
        alpha = step / num_steps                                       # path.py:139
        p = pa + alpha * (pb - pa)
        q, ok = computeqgrasppose(q, Ra, p, **kw)
        if not ok:
            break                                                      # path.py:153-156
        path.append(q)
    return path
