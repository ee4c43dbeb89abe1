"""Torch-facing wrapper of the C ABI: device memory, streams and layout conversion only -- every number is
produced by the CUDA kernels in csrc/.  Tensors must live on the handle's CUDA device; launches go to the
current torch stream and never synchronise."""
from __future__ import annotations

import ctypes
import threading
from dataclasses import dataclass

import numpy as np
import torch

from . import _cabi
from .model import KinematicTable, nextage_table
from .scene import CollisionScene, nextage_scene

# Reference literals: EPSILON (config.py:22), DT / max_iters (inverse_geometry.py:53-54)
EPSILON = 1e-3
DT = 1e-2
MAX_ITERS = 1000


@dataclass
class SolveInfo:
    iters: torch.Tensor     # [B] int32, updates applied
    resid: torch.Tensor     # [B,2] residual norms (left, right) at the returned q


def _sfx(dtype) -> str:
    if dtype == torch.float32:
        return "f32"
    if dtype == torch.float64:
        return "f64"
    raise TypeError(f"dtype must be torch.float32 or torch.float64, got {dtype}")


def as_pose12(pose, *, dtype, device, batch=None) -> torch.Tensor:
    """Normalise a cube placement batch to [B,12] (rotation row-major, translation).
    Accepts [B,12], [B,4,4], [B,7] (x y z qx qy qz qw, pinocchio XYZQUAT order) or [B,3] (identity rotation,
    the only case path.py samples: path.py:47)."""
    t = torch.as_tensor(pose, dtype=dtype, device=device) if not torch.is_tensor(pose) else pose.to(device=device, dtype=dtype)
    if t.dim() == 1 or (t.dim() == 2 and t.shape == (4, 4)):
        t = t.unsqueeze(0)
    if t.dim() == 3 and t.shape[1:] == (4, 4):
        out = torch.cat([t[:, :3, :3].reshape(-1, 9), t[:, :3, 3]], dim=1)
    elif t.dim() == 2 and t.shape[1] == 12:
        out = t
    elif t.dim() == 2 and t.shape[1] == 3:
        eye = torch.eye(3, dtype=dtype, device=device).reshape(1, 9).expand(t.shape[0], 9)
        out = torch.cat([eye, t], dim=1)
    elif t.dim() == 2 and t.shape[1] == 7:
        x, y, z, w = t[:, 3], t[:, 4], t[:, 5], t[:, 6]
        R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                         2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                         2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=1)
        out = torch.cat([R, t[:, :3]], dim=1)
    else:
        raise ValueError(f"unsupported pose shape {tuple(t.shape)}")
    if batch is not None and out.shape[0] == 1 and batch != 1:
        out = out.expand(batch, 12)
    return out


class GraspIK:
    """Handle on the flattened kinematic model living on one GPU (gik_create / gik_destroy)."""

    def __init__(self, table: KinematicTable | None = None, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("GraspIK needs a CUDA device: there is no CPU fallback")
        self.table = table if table is not None else nextage_table()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("device must be a CUDA device")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._lib = _cabi.lib()
        self._h = ctypes.c_void_p()
        ct = self.table.to_c()
        _cabi.check(self._lib.gik_create(ctypes.byref(ct), self.device.index, ctypes.byref(self._h)), "gik_create")
        self.nq = self.table.nq
        self.launches = 0   # kernels of this library launched through this handle

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.gik_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # GIK_F_LANE_KERNEL / GIK_F_PAIR_KERNEL / GIK_F_SCALAR_LANE (fp32: "lane" = packed FFMA2 kernel, "lane1" = scalar)
    _KERNEL_FLAGS = {None: 0, "auto": 0, "lane": 2, "pair": 4, "lane1": 2 | 8}

    # Step form: by default an undamped step on a spherical-wrist table (Nextage) runs as two 3x3 solves per hand
    # (gik_core.cuh "Spherical-wrist form"); `force_cholesky` (a solver attribute, or the "+chol" kernel suffix) keeps the
    # general block-Cholesky form for A/B runs and tests.  GIK_F_CHOLESKY = 32.
    force_cholesky = False

    def _params(self, eps, dt, max_iters, damping, kernel=None, early_stop=False) -> _cabi.GikParams:
        chol = self.force_cholesky
        if isinstance(kernel, str) and kernel.endswith("+chol"):
            chol, kernel = True, (kernel[:-5] or None)
        if kernel not in self._KERNEL_FLAGS:
            raise ValueError(f"kernel must be one of {list(self._KERNEL_FLAGS)} (optionally with the suffix '+chol')")
        flags = self._KERNEL_FLAGS[kernel] | (16 if early_stop else 0) | (32 if chol else 0)   # GIK_F_EARLY_STOP, GIK_F_CHOLESKY
        return _cabi.GikParams(float(eps), float(dt), float(damping), int(max_iters), flags)

    def kernel_name(self, n, dtype=torch.float32, kernel=None) -> str:
        """Kernel the batch solve launches for n problems (gik_solve_kernel_name); for reports."""
        esz = 4 if dtype == torch.float32 else 8
        flags = self._params(1e-3, 1e-2, 1, 0.0, kernel).flags
        return self._lib.gik_solve_kernel_name(self._h, esz, int(n), flags).decode()

    def _chk_dev(self, *ts):
        for t in ts:
            if t is not None and (not t.is_cuda or t.device != self.device):
                raise ValueError(f"tensor on {t.device}, handle on {self.device}")

    @staticmethod
    def _ptr(t):
        return ctypes.c_void_p(0 if t is None else t.data_ptr())

    def limits(self, dtype=torch.float64):
        return (torch.as_tensor(self.table.lower, dtype=dtype, device=self.device),
                torch.as_tensor(self.table.upper, dtype=dtype, device=self.device))

    # ------------------------------------------------------------------ SoA entry points (no copies)
    def fk_soa(self, q_soa: torch.Tensor) -> torch.Tensor:
        """q [nq][n] -> frames [2][12][n]  (gik_fk_*)."""
        self._chk_dev(q_soa)
        n = q_soa.shape[1]
        out = torch.empty((2, 12, n), dtype=q_soa.dtype, device=self.device)
        f = getattr(self._lib, f"gik_fk_{_sfx(q_soa.dtype)}")
        _cabi.check(f(self._h, n, self._ptr(q_soa.contiguous()), self._ptr(out), self._stream()), "gik_fk")
        self.launches += 1
        return out

    def jac_soa(self, q_soa: torch.Tensor) -> torch.Tensor:
        """q [nq][n] -> jac [2][6][nq][n]  (gik_jac_*)."""
        self._chk_dev(q_soa)
        n = q_soa.shape[1]
        out = torch.empty((2, 6, self.nq, n), dtype=q_soa.dtype, device=self.device)
        f = getattr(self._lib, f"gik_jac_{_sfx(q_soa.dtype)}")
        _cabi.check(f(self._h, n, self._ptr(q_soa.contiguous()), self._ptr(out), self._stream()), "gik_jac")
        self.launches += 1
        return out

    def solve_soa(self, q_init: torch.Tensor, pose: torch.Tensor, *, eps=EPSILON, dt=DT, max_iters=MAX_ITERS,
                  damping=0.0, out=None, kernel=None, early_stop=False):
        """q_init [nq][n], pose [12][n] (contiguous) -> (q [nq][n], converged u8 [n], iters i32 [n], resid [2][n]).
        `out` may carry preallocated (q, converged, iters, resid) to keep the call allocation-free.
        `kernel`: None = launcher's choice by batch size, "lane" / "pair" force a thread mapping (same results).
        `early_stop`: fast preset (GIK_F_EARLY_STOP) -- stalled problems are abandoned early; flags and converged q
        are unchanged, the q / iters of FAILED problems are not the reference's."""
        self._chk_dev(q_init, pose)
        if q_init.dtype != pose.dtype:
            raise TypeError("q_init and pose must share a dtype")
        if not (q_init.is_contiguous() and pose.is_contiguous()):
            raise ValueError("SoA inputs must be contiguous")
        n = q_init.shape[1]
        if q_init.shape[0] != self.nq or pose.shape != (12, n):
            raise ValueError(f"expected q_init [{self.nq}][n] and pose [12][n], got {tuple(q_init.shape)}, {tuple(pose.shape)}")
        if out is None:
            q = torch.empty_like(q_init)
            conv = torch.empty((n,), dtype=torch.uint8, device=self.device)
            iters = torch.empty((n,), dtype=torch.int32, device=self.device)
            resid = torch.empty((2, n), dtype=q_init.dtype, device=self.device)
        else:
            q, conv, iters, resid = out
        prm = self._params(eps, dt, max_iters, damping, kernel, early_stop)
        f = getattr(self._lib, f"gik_solve_{_sfx(q_init.dtype)}")
        _cabi.check(f(self._h, n, self._ptr(q_init), self._ptr(pose), ctypes.byref(prm), self._ptr(q),
                      self._ptr(conv), self._ptr(iters), self._ptr(resid), self._stream()), "gik_solve")
        if n:
            self.launches += 1
        return q, conv, iters, resid

    def solve_scatter_soa(self, q_init: torch.Tensor, pose: torch.Tensor, q_ptrs, conv_ptrs, n_total: int, offset: int,
                          *, eps=EPSILON, dt=DT, max_iters=MAX_ITERS, damping=0.0, kernel=None, out=None):
        """K3 fused with its all-gather (gik_solve_scatter_*): solve this rank's slab and store every result into the
        result arrays of ALL ranks.  `q_ptrs` / `conv_ptrs`: device addresses (ints) of each rank's [nq][n_total] q
        array and [n_total] flag array as mapped on THIS device (see dist.SymmetricResults).  Returns the local
        (iters [n], resid [2][n]); the caller runs the cross-rank barrier."""
        self._chk_dev(q_init, pose)
        if not (q_init.is_contiguous() and pose.is_contiguous()) or q_init.dtype != pose.dtype:
            raise ValueError("SoA inputs must be contiguous and share a dtype")
        n = q_init.shape[1]
        if q_init.shape[0] != self.nq or pose.shape != (12, n):
            raise ValueError("expected q_init [nq][n] and pose [12][n]")
        if len(q_ptrs) != len(conv_ptrs) or not q_ptrs:
            raise ValueError("need one q and one flag pointer per rank")
        if out is None:
            iters = torch.empty((n,), dtype=torch.int32, device=self.device)
            resid = torch.empty((2, n), dtype=q_init.dtype, device=self.device)
        else:
            iters, resid = out
        P = len(q_ptrs)
        qa = (ctypes.c_void_p * P)(*[int(p) for p in q_ptrs])
        ca = (ctypes.c_void_p * P)(*[int(p) for p in conv_ptrs])
        prm = self._params(eps, dt, max_iters, damping, kernel)
        f = getattr(self._lib, f"gik_solve_scatter_{_sfx(q_init.dtype)}")
        _cabi.check(f(self._h, n, self._ptr(q_init), self._ptr(pose), ctypes.byref(prm), P, qa, ca, int(n_total),
                      int(offset), self._ptr(iters), self._ptr(resid), self._stream()), "gik_solve_scatter")
        if n:
            self.launches += 2 if P > 1 else 1          # solve kernel (+ its pusher blocks) and the tail kernel
        return iters, resid

    def best_of_soa(self, q, conv, resid, n_place: int, n_restart: int):
        """Problem (p, r) at column p * n_restart + r.  -> (q_best [nq][n_place], conv u8, which i32)."""
        self._chk_dev(q, conv, resid)
        qb = torch.empty((self.nq, n_place), dtype=q.dtype, device=self.device)
        cb = torch.empty((n_place,), dtype=torch.uint8, device=self.device)
        wh = torch.empty((n_place,), dtype=torch.int32, device=self.device)
        f = getattr(self._lib, f"gik_best_of_{_sfx(q.dtype)}")
        _cabi.check(f(self._h, n_place, n_restart, self._ptr(q), self._ptr(conv), self._ptr(resid), self._ptr(qb),
                      self._ptr(cb), self._ptr(wh), self._stream()), "gik_best_of")
        if n_place:
            self.launches += 1
        return qb, cb, wh

    def project_edges_soa(self, q_start, pose_a, pose_b, num_steps, max_steps: int, *, eps=EPSILON, dt=DT,
                          max_iters=MAX_ITERS, damping=0.0, kernel=None, out=None):
        """q_start [nq][E], pose_a/b [12][E], num_steps i32 [E] -> (q_path [max_steps][nq][E], n_valid i32 [E],
        iters_total i32 [E]).  Rows >= n_valid[e] of q_path are zero."""
        self._chk_dev(q_start, pose_a, pose_b, num_steps)
        E = q_start.shape[1]
        if num_steps.dtype != torch.int32:
            raise TypeError("num_steps must be int32")
        if out is None:
            path = torch.zeros((max_steps, self.nq, E), dtype=q_start.dtype, device=self.device)
            nv = torch.empty((E,), dtype=torch.int32, device=self.device)
            itt = torch.empty((E,), dtype=torch.int32, device=self.device)
        else:
            path, nv, itt = out
        prm = self._params(eps, dt, max_iters, damping, kernel)
        f = getattr(self._lib, f"gik_project_edges_{_sfx(q_start.dtype)}")
        _cabi.check(f(self._h, E, max_steps, self._ptr(q_start.contiguous()), self._ptr(pose_a.contiguous()),
                      self._ptr(pose_b.contiguous()), self._ptr(num_steps.contiguous()), ctypes.byref(prm),
                      self._ptr(path), self._ptr(nv), self._ptr(itt), self._stream()), "gik_project_edges")
        if E:
            self.launches += 1
        return path, nv, itt

    # ------------------------------------------------------------------ collision predicate
    def attach_scene(self, scene: CollisionScene | None = None) -> "GraspIK":
        """Upload the flattened collision model (gik_scene_attach).  Default: the reference's scene."""
        self.scene = scene if scene is not None else nextage_scene()
        sc = self.scene.to_c()
        _cabi.check(self._lib.gik_scene_attach(self._h, ctypes.byref(sc)), "gik_scene_attach")
        return self

    def _need_scene(self):
        """Attach a collision scene on first use.  Only two cases are automatic: the built-in Nextage table gets the
        packaged reference scene, and a solver built from a pinocchio RobotWrapper flattens that robot's OWN
        `collision_model` (scene_from_pinocchio: its geometries, placements and pair list, whatever they are).  Any
        other table (a URDF, a hand-made KinematicTable) needs an explicit attach_scene(scene): a silently wrong
        collision model is worse than an error."""
        if getattr(self, "scene", None) is None:
            src = self.table.meta.get("source")
            robot = getattr(self, "_robot", None)
            if src == "builtin-nextage":
                self.attach_scene()
            elif robot is not None and hasattr(robot, "collision_model"):
                from .scene import scene_from_pinocchio
                self.attach_scene(scene_from_pinocchio(robot))
            else:
                raise RuntimeError("no collision scene attached: call GraspIK.attach_scene(scene) for this robot "
                                   "(scene_from_urdf / scene_from_pinocchio build one)")

    def collision_soa(self, q_soa: torch.Tensor, cube_pose_soa: torch.Tensor | None = None, sel: torch.Tensor | None = None) -> torch.Tensor:
        """tools.collision(robot, q) for every column: q [nq][n], cube_pose [12][n] (None = scene's cube placement)
        -> colliding u8 [n].  `sel` (int64 [m], device): test only those columns (gik_collision_sel_*); the other
        entries of the result are 0."""
        self._need_scene()
        self._chk_dev(q_soa, cube_pose_soa, sel)
        n = q_soa.shape[1]
        cp = None if cube_pose_soa is None else cube_pose_soa.to(q_soa.dtype).contiguous()
        if sel is not None:
            out = torch.zeros((n,), dtype=torch.uint8, device=self.device)
            if sel.numel() == 0:
                return out
            f = getattr(self._lib, f"gik_collision_sel_{_sfx(q_soa.dtype)}")
            sel = sel.to(torch.int64).contiguous()
            _cabi.check(f(self._h, n, sel.numel(), self._ptr(sel), self._ptr(q_soa.contiguous()), self._ptr(cp), self._ptr(out),
                          self._stream()), "gik_collision_sel")
            self.launches += 1
            return out
        out = torch.empty((n,), dtype=torch.uint8, device=self.device)
        f = getattr(self._lib, f"gik_collision_{_sfx(q_soa.dtype)}")
        _cabi.check(f(self._h, n, self._ptr(q_soa.contiguous()), self._ptr(cp), self._ptr(out), self._stream()), "gik_collision")
        if n:
            self.launches += 1
        return out

    def clearance_soa(self, q_soa: torch.Tensor, cube_pose_soa: torch.Tensor | None = None, threshold: float = 0.04) -> torch.Tensor:
        """distanceToObstacle(robot, q) >= threshold (tools.py:38-51, path.py:61-62) -> clear u8 [n]."""
        self._need_scene()
        self._chk_dev(q_soa, cube_pose_soa)
        n = q_soa.shape[1]
        out = torch.empty((n,), dtype=torch.uint8, device=self.device)
        f = getattr(self._lib, f"gik_clearance_{_sfx(q_soa.dtype)}")
        cp = None if cube_pose_soa is None else cube_pose_soa.to(q_soa.dtype).contiguous()
        _cabi.check(f(self._h, n, self._ptr(q_soa.contiguous()), self._ptr(cp), float(threshold), self._ptr(out),
                      self._stream()), "gik_clearance")
        if n:
            self.launches += 1
        return out

    def obstacle_distance_soa(self, q_soa: torch.Tensor, cube_pose_soa: torch.Tensor | None = None, d_max: float = 2.0) -> torch.Tensor:
        """distanceToObstacle(robot, q) (tools.py:38-51) as a value -> dist [n] in q's dtype: smallest distance between
        the robot (and cube) and the table / obstacle, capped at d_max, 0 on intersection (gik_obstacle_distance_*)."""
        self._need_scene()
        self._chk_dev(q_soa, cube_pose_soa)
        n = q_soa.shape[1]
        out = torch.empty((n,), dtype=q_soa.dtype, device=self.device)
        f = getattr(self._lib, f"gik_obstacle_distance_{_sfx(q_soa.dtype)}")
        cp = None if cube_pose_soa is None else cube_pose_soa.to(q_soa.dtype).contiguous()
        _cabi.check(f(self._h, n, self._ptr(q_soa.contiguous()), self._ptr(cp), float(d_max), self._ptr(out),
                      self._stream()), "gik_obstacle_distance")
        if n:
            self.launches += 1
        return out

    def bezier_fit(self, q0: torch.Tensor, q1: torch.Tensor, path: torch.Tensor, pinv, basis, w0, w1):
        """The least-squares fit of `maketraj` (control.py:63-195) for many paths in one launch (gik_bezier_fit_*).
        q0, q1 [N, dim], path [N, n_points, dim] on this device; pinv [n_free, n_points], basis [n_points, n_free],
        w0 / w1 [n_points]: the factored design matrix (trajectory.fit_operands).  -> (ctrl [N, n_free + 6, dim], cost [N])."""
        self._chk_dev(q0, q1, path)
        dt = path.dtype
        if dt not in (torch.float32, torch.float64) or q0.dtype != dt or q1.dtype != dt:
            raise TypeError("q0, q1 and path must share a float32 / float64 dtype")
        N, n_points, dim = path.shape
        n_free = int(pinv.shape[0])
        if q0.shape != (N, dim) or q1.shape != (N, dim) or tuple(pinv.shape) != (n_free, n_points) or \
                tuple(basis.shape) != (n_points, n_free) or w0.shape[0] != n_points or w1.shape[0] != n_points:
            raise ValueError("bezier_fit: inconsistent shapes")
        ops = [torch.as_tensor(x, dtype=dt, device=self.device).contiguous() for x in (pinv, basis, w0, w1)]
        ctrl = torch.empty((N, n_free + 6, dim), dtype=dt, device=self.device)
        cost = torch.empty((N,), dtype=dt, device=self.device)
        f = getattr(self._lib, f"gik_bezier_fit_{_sfx(dt)}")
        _cabi.check(f(self._h, N, n_points, dim, n_free + 6, *[self._ptr(x) for x in ops], self._ptr(q0.contiguous()),
                      self._ptr(q1.contiguous()), self._ptr(path.contiguous()), self._ptr(ctrl), self._ptr(cost), self._stream()),
                    "gik_bezier_fit")
        if N:
            self.launches += 1
        return ctrl, cost

    def cube_collision_soa(self, cube_pose_soa: torch.Tensor) -> torch.Tensor:
        """The cube's own collision test (cube vs table / obstacle, path.py:51-52): cube_pose [12][n] -> u8 [n]."""
        self._need_scene()
        self._chk_dev(cube_pose_soa)
        n = cube_pose_soa.shape[1]
        out = torch.empty((n,), dtype=torch.uint8, device=self.device)
        f = getattr(self._lib, f"gik_cube_collision_{_sfx(cube_pose_soa.dtype)}")
        _cabi.check(f(self._h, n, self._ptr(cube_pose_soa.contiguous()), self._ptr(out), self._stream()), "gik_cube_collision")
        if n:
            self.launches += 1
        return out

    def collision(self, q, cube_pose=None, *, dtype=None) -> torch.Tensor:
        """Row-major convenience: q [B,nq] (or [nq]), cube_pose [B,12|4x4|7|3] or None -> colliding bool [B]."""
        qt = torch.as_tensor(q, device=self.device)
        dtype = dtype or (qt.dtype if qt.dtype in (torch.float32, torch.float64) else torch.float64)
        qt = qt.to(dtype)
        if qt.dim() == 1:
            qt = qt.unsqueeze(0)
        cp = None
        if cube_pose is not None:
            cp = as_pose12(cube_pose, dtype=dtype, device=self.device, batch=qt.shape[0]).t().contiguous()
        return self.collision_soa(qt.t().contiguous(), cp).bool()

    def solve_success_soa(self, q_init, pose, *, eps=EPSILON, dt=DT, max_iters=MAX_ITERS, damping=0.0, kernel=None,
                          descend_while_colliding=True, early_stop=False, return_stats=False):
        """The reference's FULL success predicate on the device: `success = converged and not collision(q)`
        (inverse_geometry.py:70, 97-98), including its behaviour on converged-but-colliding iterates -- the loop keeps
        descending and re-tests collision(q) after every update until an iterate is free or max_iters is reached.
        ONE stream-ordered C call (gik_solve_success_*), no host synchronisation: the batched solve, the collision test
        of the converged problems, a continuation launch that runs the converged-but-colliding ones to the cap, and a
        warp-per-problem kernel that decides each of them (persistent collision, or a replay that tests the undecided
        pairs on every iterate).  `descend_while_colliding=False` (GIK_F_NO_DESCEND) evaluates the collision term once,
        at the configuration the descent stopped at: differs only where further descent would have freed a collision,
        and in the q returned for failed problems.
        -> (q [nq][n], success u8 [n], converged u8 [n], iters i32 [n], resid [2][n][, stats]); stats (device int64 [4]):
        problems decided as persistent collisions, problems replayed, replayed iterations, replays that succeeded."""
        self._need_scene()
        self._chk_dev(q_init, pose)
        if q_init.dtype != pose.dtype:
            raise TypeError("q_init and pose must share a dtype")
        n = q_init.shape[1]
        q = torch.empty_like(q_init)
        succ = torch.empty((n,), dtype=torch.uint8, device=self.device)
        conv = torch.empty((n,), dtype=torch.uint8, device=self.device)
        iters = torch.empty((n,), dtype=torch.int32, device=self.device)
        resid = torch.empty((2, n), dtype=q_init.dtype, device=self.device)
        esz = 4 if q_init.dtype == torch.float32 else 8
        nbytes = int(self._lib.gik_solve_success_scratch_bytes(self._h, n, esz))
        scratch = torch.empty((max(nbytes, 8),), dtype=torch.uint8, device=self.device)
        prm = self._params(eps, dt, max_iters, damping, kernel, early_stop)
        if not descend_while_colliding:
            prm.flags |= 64                                                  # GIK_F_NO_DESCEND
        f = getattr(self._lib, f"gik_solve_success_{_sfx(q_init.dtype)}")
        _cabi.check(f(self._h, n, self._ptr(q_init.contiguous()), self._ptr(pose.contiguous()), ctypes.byref(prm),
                      self._ptr(q), self._ptr(succ), self._ptr(conv), self._ptr(iters), self._ptr(resid),
                      self._ptr(scratch), self._stream()), "gik_solve_success")
        if n:
            self.launches += 3 if not descend_while_colliding else 6
        if return_stats:
            stats = scratch[nbytes - 256 + 8:nbytes - 256 + 40].view(torch.int64).clone() if n else torch.zeros(4, dtype=torch.int64, device=self.device)
            return q, succ, conv, iters, resid, stats
        return q, succ, conv, iters, resid

    # ------------------------------------------------------------------ row-major convenience ([B, ...])
    def fk(self, q: torch.Tensor):
        """q [B,nq] -> (R [B,2,3,3], p [B,2,3]) world placements of LARM_EFF / RARM_EFF."""
        fr = self.fk_soa(q.t().contiguous())              # [2][12][B]
        fr = fr.permute(2, 0, 1)                          # [B,2,12]
        return fr[:, :, :9].reshape(-1, 2, 3, 3), fr[:, :, 9:]

    def jacobians(self, q: torch.Tensor):
        """q [B,nq] -> J [B,2,6,nq] (LOCAL frame, rows [linear; angular])."""
        return self.jac_soa(q.t().contiguous()).permute(3, 0, 1, 2)

    def solve(self, q_init, pose, *, dtype=torch.float32, eps=EPSILON, dt=DT, max_iters=MAX_ITERS, damping=0.0,
              return_info=False, kernel=None):
        """q_init [B,nq] (or [nq], broadcast), pose [B,12|4x4|7|3] -> (q [B,nq], converged bool [B][, SolveInfo])."""
        p12 = as_pose12(pose, dtype=dtype, device=self.device)
        B = p12.shape[0]
        qi = torch.as_tensor(q_init, device=self.device).to(dtype)
        if qi.dim() == 1:
            qi = qi.unsqueeze(0)
        if qi.shape[0] == 1 and B != 1:
            qi = qi.expand(B, self.nq)
        elif B == 1 and qi.shape[0] != 1:
            B = qi.shape[0]
            p12 = p12.expand(B, 12)
        q, conv, iters, resid = self.solve_soa(qi.t().contiguous(), p12.t().contiguous(), eps=eps, dt=dt,
                                               max_iters=max_iters, damping=damping, kernel=kernel)
        res = (q.t(), conv.bool())
        if return_info:
            res = res + (SolveInfo(iters, resid.t()),)
        return res


def _pinned(cache, key, shape, dtype):
    t = cache.get(key)
    if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
        t = torch.empty(shape, dtype=dtype).pin_memory()
        cache[key] = t
    return t


def _slab_bounds(B: int, slabs: int, align: int = 32):
    """Slab boundaries of the streamed host path: a small first slab so the kernel starts early, the rest in equal
    parts, every interior boundary a multiple of `align` rows (rows are 60 / 48 B: a multiple of 32 rows is a multiple
    of 128 B, so no cache line ever holds rows of two slabs)."""
    slabs = max(1, int(slabs))
    if slabs == 1 or B <= (1 << 16):                 # small batches: one slab, nothing to overlap
        return [0, B]
    first = min(B, max(1 << 16, B // 16))
    first -= first % align
    if first <= 0 or first >= B:
        return [0, B]
    bounds = [0, first]
    for k in range(1, slabs):
        b = first + ((B - first) * k) // (slabs - 1)
        if k < slabs - 1:
            b -= b % align
        if b > bounds[-1]:
            bounds.append(b)
    bounds[-1] = B
    return bounds


def solve_host(self, q_init, pose, *, dtype=torch.float32, eps=EPSILON, dt=DT, max_iters=MAX_ITERS, damping=0.0,
               slabs=4, return_info=False, kernel=None, out=None):
    """Host-buffer entry: q_init [B,nq] (or [nq]) and pose [B,12|4x4|7|3] are CPU tensors / arrays (pageable or pinned);
    returns CPU tensors (q [B,nq], converged bool [B][, SolveInfo]).

    Ownership: by default the results are FRESH pinned tensors owned by the caller (one pinned allocation per call).
    `out=(q [B,nq], converged uint8 [B][, iters int32 [B], resid [B,2]])` -- pinned, contiguous CPU tensors -- makes the
    call allocation-free: the kernel stores straight into them (the zero-copy path bench.py's e2e leg uses).  The
    device staging buffers are cached on the solver and guarded by a lock: concurrent calls on one solver serialise.

    ONE solve launch (gik_solve_rows_*), overlapped with its own transfers:
      * inputs: the copy engine streams the row-major host arrays into device staging buffers slab by slab on a copy
        stream and advances a device counter after each slab; the kernel is launched as soon as the first (small) slab
        is in, and a lane refill only waits if it runs ahead of the copies (copies deliver ~20x faster than the solver
        consumes, so in practice never);
      * outputs: the kernel stores q / converged (/ iters / resid) of each finished problem straight into the pinned
        host arrays (UVA stores over PCIe), so no device-to-host copy follows the kernel."""
    dev = self.device
    p12 = as_pose12(pose, dtype=dtype, device="cpu")
    B = p12.shape[0]
    qi = torch.as_tensor(q_init).to(dtype)
    if qi.dim() == 1:
        qi = qi.unsqueeze(0)
    if qi.shape[0] == 1 and B != 1:
        qi = qi.expand(B, self.nq)
    if out is not None:
        q_out, c_out = out[0], out[1]
        it_out = out[2] if len(out) > 2 else None
        r_out = out[3] if len(out) > 3 else None
        if return_info and (it_out is None or r_out is None):
            raise ValueError("return_info=True needs out=(q, converged, iters, resid)")
        want = [(q_out, (B, self.nq), dtype), (c_out, (B,), torch.uint8)]
        if it_out is not None:
            want.append((it_out, (B,), torch.int32))
        if r_out is not None:
            want.append((r_out, (B, 2), dtype))
        for t, shape, dt_ in want:
            if tuple(t.shape) != shape or t.dtype != dt_ or not t.is_contiguous() or t.is_cuda or (B and not t.is_pinned()):
                raise ValueError(f"out tensors must be pinned contiguous CPU tensors; expected {shape} {dt_}")
    else:
        q_out = torch.empty((B, self.nq), dtype=dtype, pin_memory=True) if B else torch.empty((0, self.nq), dtype=dtype)
        c_out = torch.empty((B,), dtype=torch.uint8, pin_memory=True) if B else torch.empty((0,), dtype=torch.uint8)
        it_out = (torch.empty((B,), dtype=torch.int32, pin_memory=True) if B else torch.empty((0,), dtype=torch.int32)) if return_info else None
        r_out = (torch.empty((B, 2), dtype=dtype, pin_memory=True) if B else torch.empty((0, 2), dtype=dtype)) if return_info else None
    if B == 0:
        return (q_out, c_out.view(torch.bool)) + ((SolveInfo(it_out, r_out),) if return_info else ())
    lock = self.__dict__.setdefault("_host_lock", threading.Lock())
    with lock:
        cache = self.__dict__.setdefault("_host_cache", {})
        # pageable inputs are staged through cached pinned buffers (a host memcpy); pinned inputs are used in place
        q_in = qi if (qi.is_pinned() and qi.is_contiguous()) else _pinned(cache, "q_in", (B, self.nq), dtype).copy_(qi)
        p_in = p12 if (p12.is_pinned() and p12.is_contiguous()) else _pinned(cache, "p_in", (B, 12), dtype).copy_(p12)
        key = (B, dtype)
        if cache.get("stage_key") != key:
            cache["stage_key"] = key
            cache["dq"] = torch.empty((B, self.nq), dtype=dtype, device=dev)
            cache["dp"] = torch.empty((B, 12), dtype=dtype, device=dev)
            cache["ready"] = torch.zeros((1,), dtype=torch.int64, device=dev)
        dq, dp, ready = cache["dq"], cache["dp"], cache["ready"]
        if self.__dict__.get("_poison_staging"):          # test hook: stale staging contents must never be read
            dq.fill_(float("nan")); dp.fill_(float("nan"))
        copy_stream = cache.get("copy_stream")
        if copy_stream is None:
            copy_stream = cache["copy_stream"] = torch.cuda.Stream(device=dev)
        bounds = _slab_bounds(B, slabs)
        marks = _pinned(cache, "marks", (len(bounds) - 1,), torch.int64)
        marks.copy_(torch.tensor(bounds[1:], dtype=torch.int64))
        cur = torch.cuda.current_stream(dev)
        copy_stream.wait_stream(cur)                      # staging buffers may still be read by a previous launch
        first_in = torch.cuda.Event()
        with torch.cuda.stream(copy_stream):
            ready.zero_()
            for k in range(len(bounds) - 1):
                lo, hi = bounds[k], bounds[k + 1]
                dq[lo:hi].copy_(q_in[lo:hi], non_blocking=True)
                dp[lo:hi].copy_(p_in[lo:hi], non_blocking=True)
                ready.copy_(marks[k:k + 1], non_blocking=True)
                if k == 0:
                    first_in.record(copy_stream)
        cur.wait_event(first_in)
        prm = self._params(eps, dt, max_iters, damping, kernel)
        f = getattr(self._lib, f"gik_solve_rows_{_sfx(dtype)}")
        _cabi.check(f(self._h, B, self._ptr(dq), self._ptr(dp), ctypes.byref(prm), self._ptr(q_out), self._ptr(c_out),
                      self._ptr(it_out), self._ptr(r_out), self._ptr(ready), self._stream()), "gik_solve_rows")
        self.launches += 1
        cur.synchronize()          # results are host memory: the call returns when the kernel has stored them
        copy_stream.synchronize()
    return (q_out, c_out.view(torch.bool)) + ((SolveInfo(it_out, r_out),) if return_info else ())


GraspIK.solve_host = solve_host

_default = {}


def default_solver(device=None) -> GraspIK:
    """Process-wide GraspIK for the built-in Nextage table on `device` (current CUDA device by default)."""
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: the batched grasp IK has no CPU fallback")
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    if idx not in _default:
        _default[idx] = GraspIK(nextage_table(), torch.device("cuda", idx))
    return _default[idx]


def fma_peak_tflops(device_index: int, elem_size: int, repeats: int = 5) -> float:
    out = ctypes.c_double(0.0)
    _cabi.check(_cabi.lib().gik_measure_fma_peak(device_index, elem_size, repeats, ctypes.byref(out)), "gik_measure_fma_peak")
    return out.value


def flops_per_iter() -> int:
    return int(_cabi.lib().gik_flops_per_iter())


def flops_per_iter_executed(elem_size: int, wrist: bool) -> int:
    return int(_cabi.lib().gik_flops_per_iter_executed(int(elem_size), 1 if wrist else 0))


def bytes_per_solve(elem_size: int) -> int:
    return int(_cabi.lib().gik_bytes_per_solve(elem_size))
