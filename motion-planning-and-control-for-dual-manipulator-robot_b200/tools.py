"""The helpers of the reference's `tools.py` that sit on the IK path, with the same names and call shapes, backed by
the flattened model / scene and the CUDA kernels instead of pinocchio + hpp-fcl:

    jointlimitscost, jointlimitsviolated, projecttojointlimits   (tools.py:11-22)
    collision(robot, q)                                          (tools.py:25-35)   -> gik_collision_*
    distanceToObstacle(robot, q)                                 (tools.py:38-51)   -> gik_obstacle_distance_*
    getcubeplacement / setcubeplacement                          (tools.py:54-68)   -> the solver's current cube pose

`robot` is whatever `solver_for` accepts (pinocchio RobotWrapper, KinematicTable, GraspIK, None = built-in Nextage).
The cube placement that the reference stores inside the pinocchio geometry objects lives on the solver here
(`solver.cube_pose`, a 12-vector); `setcubeplacement` updates it, `collision` / `distanceToObstacle` use it."""
from __future__ import annotations

import numpy as np
import torch

from .inverse_geometry import _pose_to_array, _setcubeplacement, solver_for


def _limits(robot):
    """Joint limits without touching the GPU: from the table, the pinocchio model or the built-in Nextage constants."""
    from .model import KinematicTable, nextage_table
    from .ops import GraspIK
    if isinstance(robot, GraspIK):
        t = robot.table
    elif isinstance(robot, KinematicTable):
        t = robot
    elif robot is None:
        t = nextage_table()
    else:
        m = getattr(robot, "model", robot)
        return np.asarray(m.lowerPositionLimit, float).reshape(-1), np.asarray(m.upperPositionLimit, float).reshape(-1)
    return np.asarray(t.lower, float), np.asarray(t.upper, float)


def jointlimitscost(robot, q):
    lo, hi = _limits(robot)
    q = np.asarray(q, float)
    return max(0.0, max(float(np.max(q - hi)), float(np.max(lo - q))))


def jointlimitsviolated(robot, q):
    """Return true if config not in joint limits (tools.py:17-19)."""
    return jointlimitscost(robot, q) > 0.0


def projecttojointlimits(robot, q):
    lo, hi = _limits(robot)
    return np.minimum(np.maximum(lo, np.asarray(q, float)), hi)


def setcubeplacement(robot, cube, oMf):
    """tools.setcubeplacement (tools.py:62-68): remember the cube placement for the following collision queries (and
    write it into the pinocchio geometry objects when real pinocchio wrappers are passed)."""
    s = solver_for(robot, cube)
    s.cube_pose = _pose_to_array(oMf)
    _setcubeplacement(robot, cube, oMf)


def _is_cube_wrapper(obj):
    """A pinocchio wrapper of the reference's cube (cube_small.urdf: one collision geometry, LARM_HOOK / RARM_HOOK frames)."""
    m, cm = getattr(obj, "model", None), getattr(obj, "collision_model", None)
    try:
        return m is not None and cm is not None and hasattr(m, "existFrame") and bool(m.existFrame("LARM_HOOK"))
    except Exception:
        return False


def getcubeplacement(robot, hookname=None):
    """tools.getcubeplacement (tools.py:54-59).  With the reference's own argument -- the CUBE wrapper -- the placement is
    read from its pinocchio geometry object exactly as the reference does and returned as a pinocchio SE3; with a robot /
    solver / None it is the solver's current cube pose as a 12-vector (rotation row-major, translation).  With a hook
    name ('LARM_HOOK' / 'RARM_HOOK') the placement of that hook frame, cube * hook offset."""
    if _is_cube_wrapper(robot):
        cube = robot
        oMf = cube.collision_model.geometryObjects[0].placement
        if hookname is not None:
            oMf = oMf * cube.data.oMf[cube.model.getFrameId(hookname)]
        return oMf
    s = solver_for(robot)
    pose = getattr(s, "cube_pose", None)
    if pose is None:
        s._need_scene()
        g = s.scene.geoms[s.scene.cube]
        pose = np.concatenate([np.asarray(g.R, float).reshape(9), np.asarray(g.p, float)])
    if hookname is None:
        return pose.copy()
    h = {"LARM_HOOK": 0, "RARM_HOOK": 1}[hookname]
    R, p = pose[:9].reshape(3, 3), pose[9:]
    return np.concatenate([(R @ s.table.hook_R[h]).reshape(9), p + R @ s.table.hook_p[h]])


def _q_soa(s, q, dtype):
    return torch.as_tensor(np.asarray(q, float), dtype=dtype, device=s.device).reshape(1, -1).t().contiguous()


def _cube_soa(s, dtype):
    pose = getattr(s, "cube_pose", None)
    return None if pose is None else torch.as_tensor(pose, dtype=dtype, device=s.device).reshape(1, 12).t().contiguous()


def collision(robot, q, dtype=torch.float64, cube=None):
    """Return true if in collision, false otherwise (tools.py:25-35).  `cube`: the cube wrapper, needed only the first
    time a pinocchio robot is seen (the solver flattens robot + cube once and is cached afterwards)."""
    s = solver_for(robot, cube)
    return bool(s.collision_soa(_q_soa(s, q, dtype), _cube_soa(s, dtype))[0].item())


def distanceToObstacle(robot, q, dtype=torch.float64, d_max=2.0, cube=None):
    """Shortest distance between the robot and the obstacle / table (tools.py:38-51): ONE launch of
    gik_obstacle_distance_* (each pair's distance bracketed by bisection on the margin of the kernels' boolean GJK, to
    d_max / 2^40 in float64) and one read-back.  Returns 0.0 when a pair intersects: hpp-fcl reports the negative
    penetration depth there, and every caller in the reference only compares the value with a positive threshold
    (path.py:61-62), so the sign below zero is not reproduced.  Capped at `d_max`."""
    s = solver_for(robot, cube)
    s._need_scene()
    return float(s.obstacle_distance_soa(_q_soa(s, q, dtype), _cube_soa(s, dtype), d_max)[0].item())
