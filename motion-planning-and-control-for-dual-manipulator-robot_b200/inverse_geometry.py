"""Drop-in for the reference's `inverse_geometry.py`: same name, same signature, same return value
(`computeqgrasppose(robot, qcurrent, cube, cubetarget, viz=None) -> (q, success)`, inverse_geometry.py:17,100),
plus the new batched entry point `computeqgrasppose_batch`.  All arithmetic runs in the CUDA kernels behind
include/gik.h; this file only adapts arguments and reproduces the reference's side effects.

What is preserved from the reference function
  * `qcurrent` is not modified (copy at inverse_geometry.py:49); q is returned as float64 ndarray [nq].
  * `setcubeplacement(robot, cube, cubetarget)` side effect (inverse_geometry.py:42 -> tools.py:62-68) when
    `robot` / `cube` are pinocchio wrappers, so a later `collision(robot, q)` by the caller sees the cube there.
  * `success` = the loop broke with both residuals < EPSILON and `not collision(robot, q)` (:70), and the final
    configuration does not collide (:97-98).  While converged-but-colliding the reference keeps descending; that
    is reproduced by single-step re-entry of the kernel with the collision test in between.
  * never raises on non-convergence.
`collision` is taken from (in this order) the `collision=` keyword (a callable q -> bool, or None to skip the test),
the reference's own pinocchio `computeCollisions` when pinocchio objects are passed, else the GPU collision kernel on
the packaged reference scene (scene.nextage_scene(): same geometries, pairs and placements as setup_pinocchio.py
builds) when the solver runs the built-in Nextage table.
"""
from __future__ import annotations

import weakref

import numpy as np
import torch

from .model import KinematicTable, from_pinocchio
from .ops import DT, EPSILON, MAX_ITERS, GraspIK, as_pose12, default_solver

_solvers = weakref.WeakKeyDictionary()
_solvers_by_id = {}          # objects that cannot be weakly referenced: id -> (object kept alive, solver)


def _is_pinocchio_wrapper(robot) -> bool:
    return hasattr(robot, "model") and hasattr(robot.model, "jointPlacements")


def solver_for(robot, cube=None, device=None) -> GraspIK:
    """GraspIK for `robot`: a GraspIK is returned as is; a KinematicTable is flattened once; a pinocchio
    RobotWrapper is flattened with `from_pinocchio` (cached per object); None -> built-in Nextage table."""
    if isinstance(robot, GraspIK):
        return robot
    if robot is None:
        return default_solver(device)
    if isinstance(robot, KinematicTable):
        key = robot
    elif _is_pinocchio_wrapper(robot):
        key = robot
    else:
        raise TypeError("robot must be a pinocchio RobotWrapper, a KinematicTable, a GraspIK or None")
    try:
        return _solvers[key]
    except (KeyError, TypeError):
        pass
    hit = _solvers_by_id.get(id(key))
    if hit is not None and hit[0] is key:
        return hit[1]
    if not isinstance(robot, KinematicTable) and cube is None:
        raise ValueError("the first call for a pinocchio robot needs `cube` (its LARM_HOOK / RARM_HOOK frames are part of "
                         "the flattened table)")
    table = robot if isinstance(robot, KinematicTable) else from_pinocchio(robot, cube)
    s = GraspIK(table, device)
    if not isinstance(robot, KinematicTable):
        s._robot = robot          # its collision_model is flattened on the first collision query (GraspIK._need_scene)
    try:
        _solvers[key] = s
    except TypeError:
        _solvers_by_id[id(key)] = (key, s)
    return s


def _pose_to_array(cubetarget) -> np.ndarray:
    """pin.SE3 (has .rotation/.translation), (R, p) tuple, 4x4, 12-, 7- or 3-vector -> [12] float64."""
    if hasattr(cubetarget, "rotation") and hasattr(cubetarget, "translation"):
        return np.concatenate([np.asarray(cubetarget.rotation, float).reshape(9),
                               np.asarray(cubetarget.translation, float).reshape(3)])
    if isinstance(cubetarget, (tuple, list)) and len(cubetarget) == 2 and np.ndim(cubetarget[0]) == 2:
        return np.concatenate([np.asarray(cubetarget[0], float).reshape(9), np.asarray(cubetarget[1], float).reshape(3)])
    a = np.asarray(cubetarget, float)
    return as_pose12(torch.from_numpy(a), dtype=torch.float64, device="cpu")[0].numpy()


def _setcubeplacement(robot, cube, cubetarget) -> None:
    """tools.setcubeplacement (tools.py:62-68) for pinocchio objects; no-op otherwise."""
    if not (_is_pinocchio_wrapper(robot) and hasattr(cube, "collision_model")):
        return
    try:
        import pinocchio as pin
    except ImportError:
        return
    oMf = cubetarget
    if not (hasattr(oMf, "rotation") and hasattr(oMf, "translation")):
        a = _pose_to_array(cubetarget)
        oMf = pin.SE3(a[:9].reshape(3, 3), a[9:])
    robot.visual_model.geometryObjects[-1].placement = oMf
    robot.collision_model.geometryObjects[-1].placement = oMf
    cube.visual_model.geometryObjects[-1].placement = oMf
    cube.collision_model.geometryObjects[0].placement = oMf
    pin.updateGeometryPlacements(cube.model, cube.data, cube.collision_model, cube.collision_data, cube.q0)


def _reference_collision(robot):
    """The reference's tools.collision(robot, q) (tools.py:25-35), restated on the pinocchio API."""
    if not (_is_pinocchio_wrapper(robot) and hasattr(robot, "collision_model")):
        return None
    try:
        import pinocchio as pin
    except ImportError:
        return None

    def collision(q):
        pin.updateGeometryPlacements(robot.model, robot.data, robot.collision_model, robot.collision_data, q)
        return pin.computeCollisions(robot.collision_model, robot.collision_data, False)

    return collision


def computeqgrasppose(robot, qcurrent, cube, cubetarget, viz=None, *, collision="auto", dtype=torch.float64,
                      eps=EPSILON, dt=DT, max_iters=MAX_ITERS, damping=0.0, return_info=False):
    """Reference-compatible single solve (inverse_geometry.py:17-100).  Runs on the GPU in `dtype`
    (float64 by default so results match the reference at round-off level)."""
    solver = solver_for(robot, cube)
    _setcubeplacement(robot, cube, cubetarget)
    pose = torch.from_numpy(_pose_to_array(cubetarget)).to(solver.device)
    q0 = torch.as_tensor(np.asarray(qcurrent, dtype=np.float64).copy(), device=solver.device)
    device_scene = False
    if collision == "auto":
        collision = _reference_collision(robot)
        if collision is None:
            try:
                solver._need_scene()
                device_scene = True
            except RuntimeError:
                collision = None          # no collision model known for this robot: success = converged (documented)

    if device_scene:
        # the whole predicate on the device, one call and one read-back: descent, collision test, and the reference's
        # keep-descending-while-colliding tail (gik_solve_success_*)
        qs, succ, _, it_d, res_d = solver.solve_success_soa(q0.to(dtype).unsqueeze(1).contiguous(),
                                                           pose.to(dtype).unsqueeze(1).contiguous(), eps=eps, dt=dt,
                                                           max_iters=max_iters, damping=damping)
        q_np = qs[:, 0].double().cpu().numpy().copy()
        success = bool(succ[0].item())
        iters = int(it_d[0].item())
        resid = res_d[:, 0].double().cpu().numpy()
    else:
        q, conv, info = solver.solve(q0, pose, dtype=dtype, eps=eps, dt=dt, max_iters=max_iters, damping=damping,
                                     return_info=True)
        q_np = q[0].double().cpu().numpy().copy()
        success = bool(conv[0].item())
        iters = int(info.iters[0].item())
        resid = info.resid[0].double().cpu().numpy()

        if success and collision is not None and collision(q_np):
            # converged but colliding under a HOST predicate (the caller's callable, or pinocchio's own computeCollisions):
            # the reference keeps iterating (predicate at :70 stays False).  Re-enter the kernel one update at a time
            # (eps below any reachable residual forces the step), testing in between.
            success = False
            tiny = float(np.finfo(np.float32 if dtype == torch.float32 else np.float64).tiny)
            while iters < max_iters:
                qd, _, inf1 = solver.solve(torch.from_numpy(q_np).to(solver.device), pose, dtype=dtype, eps=tiny, dt=dt,
                                           max_iters=1, damping=damping, return_info=True)
                q_np = qd[0].double().cpu().numpy().copy()
                resid = inf1.resid[0].double().cpu().numpy()
                iters += 1
                if iters < max_iters and resid[0] < eps and resid[1] < eps and not collision(q_np):
                    success = True
                    break
        if collision is not None and collision(q_np):   # inverse_geometry.py:97-98
            success = False

    if viz is not None:
        try:
            from setup_meshcat import updatevisuals   # reference helper, when its modules are importable
            updatevisuals(viz, robot, cube, q_np)
        except ImportError:
            if hasattr(viz, "display"):
                viz.display(q_np)
    if return_info:
        return q_np, success, iters, resid
    return q_np, success


def computeqgrasppose_batch(robot, q_init, cube_pose, *, dtype=torch.float32, eps=EPSILON, dt=DT,
                            max_iters=MAX_ITERS, damping=0.0, restarts=1, generator=None, return_info=False,
                            collision=False, cube=None, out=None):
    """Batched entry point (new).  `robot`: pinocchio RobotWrapper / KinematicTable / GraspIK / None (Nextage).
    q_init [B,nq] or [nq]; cube_pose [B,12|4x4|7|3].  CUDA tensors in -> CUDA tensors out (q [B,nq], converged bool
    [B][, SolveInfo]) with no host synchronisation; CPU tensors / numpy arrays in -> CPU tensors out through the pipelined
    host path (GraspIK.solve_host: H2D, solve and D2H overlap in one launch).  The returned host tensors are freshly
    allocated and owned by the caller; `out=(q, converged_u8[, iters, resid])` (pinned CPU tensors) makes the host path
    allocation-free by storing into the caller's buffers instead.  With restarts=R > 1, restart 0 starts from q_init and restarts
    1..R-1 from configurations drawn uniformly inside the joint limits; the best converged candidate (smallest
    max residual, ties -> lowest restart) is returned (BASELINE config 3; with `collision` the selection runs over the
    candidates whose full predicate holds).  `collision=True` makes the returned flag
    the reference's full `success` (converged and collision-free on the attached scene, with the reference's
    keep-descending-while-colliding behaviour: GraspIK.solve_success_soa); `collision="once"` evaluates the collision
    term once at the configuration the descent stopped at (one sync-free call, gik_solve_success_*: same decisions
    unless further descent would have freed a collision); the default returns `converged` only (the north_star's
    kernel contract; `apply_collision` applies a host-side test afterwards)."""
    solver = solver_for(robot, cube)
    host_in = (not torch.is_tensor(cube_pose) or not cube_pose.is_cuda) and (not torch.is_tensor(q_init) or not q_init.is_cuda)
    if out is not None and not (host_in and not collision and restarts <= 1):
        raise ValueError("out= is only supported on the host path (CPU inputs, no collision term, restarts=1)")
    if host_in and not collision and restarts <= 1:
        # CPU tensors in -> CPU (pinned) tensors out, copies pipelined with the solve (GraspIK.solve_host)
        return solver.solve_host(q_init, cube_pose, dtype=dtype, eps=eps, dt=dt, max_iters=max_iters, damping=damping,
                                 return_info=return_info, out=out)
    if collision and restarts <= 1:
        p12 = as_pose12(cube_pose, dtype=dtype, device=solver.device)
        B = p12.shape[0]
        qi = torch.as_tensor(q_init, device=solver.device).to(dtype)
        if qi.dim() == 1:
            qi = qi.unsqueeze(0).expand(B, solver.nq)
        q, succ, conv, iters, resid = solver.solve_success_soa(qi.t().contiguous(), p12.t().contiguous(), eps=eps, dt=dt,
                                                              max_iters=max_iters, damping=damping,
                                                              descend_while_colliding=(collision != "once"))
        res = (q.t(), succ.bool())
        if return_info:
            from .ops import SolveInfo
            info = SolveInfo(iters, resid.t())
            info.converged = conv.bool()
            res = res + (info,)
        return res
    if restarts <= 1:
        return solver.solve(q_init, cube_pose, dtype=dtype, eps=eps, dt=dt, max_iters=max_iters, damping=damping,
                            return_info=return_info)
    dev = solver.device
    p12 = as_pose12(cube_pose, dtype=dtype, device=dev)
    B = p12.shape[0]
    qi = torch.as_tensor(q_init, device=dev).to(dtype)
    if qi.dim() == 1:
        qi = qi.unsqueeze(0).expand(B, solver.nq)
    lo, hi = solver.limits(dtype)
    u = torch.rand((B, restarts, solver.nq), dtype=dtype, device=dev, generator=generator)
    cand = lo + u * (hi - lo)
    cand[:, 0, :] = qi
    # column index p * R + r, SoA
    q_soa = cand.reshape(B * restarts, solver.nq).t().contiguous()
    pose_soa = p12.unsqueeze(1).expand(B, restarts, 12).reshape(B * restarts, 12).t().contiguous()
    if collision:
        # every candidate gets the full predicate (collision term + keep-descending tail on the device); the selection
        # then runs over the SUCCESSFUL candidates, so a colliding restart never shadows a free one
        q, conv, _, iters, resid = solver.solve_success_soa(q_soa, pose_soa, eps=eps, dt=dt, max_iters=max_iters,
                                                           damping=damping, descend_while_colliding=(collision != "once"))
    else:
        q, conv, iters, resid = solver.solve_soa(q_soa, pose_soa, eps=eps, dt=dt, max_iters=max_iters, damping=damping)
    qb, cb, wh = solver.best_of_soa(q, conv, resid, B, restarts)
    res = (qb.t(), cb.bool())
    if return_info:
        from .ops import SolveInfo
        sel = torch.arange(B, device=dev) * restarts + wh.long()
        info = SolveInfo(iters[sel], resid.t()[sel])
        info.which = wh
        res = res + (info,)
    return res


def apply_collision(q, converged, collision) -> torch.Tensor:
    """success = converged and not collision(q), with the caller's host-side `collision(q: ndarray) -> bool`
    (the reference's tools.collision, tools.py:25-35) evaluated only where converged."""
    qc = q.detach().double().cpu().numpy()
    ok = converged.detach().cpu().numpy().astype(bool).copy()
    for i in np.nonzero(ok)[0]:
        if collision(qc[i]):
            ok[i] = False
    return torch.from_numpy(ok).to(converged.device)
