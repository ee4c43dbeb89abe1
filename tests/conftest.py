import ctypes
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return json.load(open(os.path.join(GOLDEN, "ik_golden.json")))


@pytest.fixture(scope="session")
def table():
    import gik_b200
    return gik_b200.nextage_table()


@pytest.fixture(scope="session")
def table_c(table):
    return table.to_c()


@pytest.fixture(scope="session")
def c_oracle():
    from oracle import c_oracle as co
    co.build()
    return co


class HostSim:
    """The device arithmetic (csrc/gik_core.cuh) compiled for the host -- test-only, see tests/hostsim."""

    def __init__(self):
        src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
        out = os.path.join(ROOT, "tests", "hostsim", "_build", "libhostsim.so")
        deps = [src] + [os.path.join(ROOT, "motion-planning-and-control-for-dual-manipulator-robot_b200", "csrc", f)
                        for f in ("gik_core.cuh", "gik_table.h", "gik_collide.cuh")] + [os.path.join(ROOT, "include", "gik.h")]
        if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
            os.makedirs(os.path.dirname(out), exist_ok=True)
            subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src], check=True)
        self.lib = ctypes.CDLL(out)

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(ctypes.c_void_p)

    def fk(self, tc, q_rows, dtype):
        from gik_b200._cabi import GikParams  # noqa: F401
        q = np.ascontiguousarray(np.asarray(q_rows, dtype).T)
        n = q.shape[1]
        fr = np.zeros((2, 12, n), dtype)
        f = self.lib.hostsim_fk_f32 if dtype == np.float32 else self.lib.hostsim_fk_f64
        assert f(ctypes.byref(tc), ctypes.c_int64(n), self._p(q), self._p(fr)) == 0
        fr = fr.transpose(2, 0, 1)
        return fr[:, :, :9].reshape(n, 2, 3, 3), fr[:, :, 9:]

    def jac(self, tc, q_rows, dtype):
        q = np.ascontiguousarray(np.asarray(q_rows, dtype).T)
        nq, n = q.shape
        J = np.zeros((2, 6, nq, n), dtype)
        f = self.lib.hostsim_jac_f32 if dtype == np.float32 else self.lib.hostsim_jac_f64
        assert f(ctypes.byref(tc), ctypes.c_int64(n), self._p(q), self._p(J)) == 0
        return J.transpose(3, 0, 1, 2)

    def solve(self, tc, q_rows, pose_rows, dtype, eps=1e-3, dt=1e-2, max_iters=1000, damping=0.0, flags=0):
        from gik_b200._cabi import GikParams
        q0 = np.ascontiguousarray(np.asarray(q_rows, dtype).T)
        pose = np.ascontiguousarray(np.asarray(pose_rows, dtype).T)
        nq, n = q0.shape
        q = np.zeros((nq, n), dtype); c = np.zeros(n, np.uint8); it = np.zeros(n, np.int32); r = np.zeros((2, n), dtype)
        prm = GikParams(eps, dt, damping, max_iters, flags)
        f = self.lib.hostsim_solve_f32 if dtype == np.float32 else self.lib.hostsim_solve_f64
        assert f(ctypes.byref(tc), ctypes.c_int64(n), self._p(q0), self._p(pose), ctypes.byref(prm), self._p(q),
                 self._p(c), self._p(it), self._p(r)) == 0
        return q.T, c.astype(bool), it, r.T


    def atan2_pos(self, y, x, dtype):
        y = np.ascontiguousarray(np.asarray(y, dtype)); x = np.ascontiguousarray(np.asarray(x, dtype))
        out = np.zeros_like(y)
        f = self.lib.hostsim_atan2_pos_f32 if dtype == np.float32 else self.lib.hostsim_atan2_pos_f64
        f.restype = None
        f(ctypes.c_int64(y.size), self._p(y), self._p(x), self._p(out))
        return out

    def collide(self, tc, sc, q_rows, cube_rows, dtype, mode=0, margin=0.0):
        """mode 0: collision(q); 1: some table/obstacle pair closer than `margin`; 2: cube vs table/obstacle."""
        f = self.lib.hostsim_collide_f32 if dtype == np.float32 else self.lib.hostsim_collide_f64
        q = None if q_rows is None else np.ascontiguousarray(np.asarray(q_rows, dtype).T)
        cube = None if cube_rows is None else np.ascontiguousarray(np.asarray(cube_rows, dtype).T)
        n = q.shape[1] if q is not None else cube.shape[1]
        out = np.zeros(n, np.uint8)
        rc = f(ctypes.byref(tc), ctypes.byref(sc), ctypes.c_int64(n), None if q is None else self._p(q),
               None if cube is None else self._p(cube), ctypes.c_int(mode), ctypes.c_double(margin), self._p(out))
        assert rc == 0, rc
        return out.astype(bool)

    def pair(self, ta, Ra, pa, sa, tb, Rb, pb, sb, dtype, margin=0.0):
        f = self.lib.hostsim_pair_f32 if dtype == np.float32 else self.lib.hostsim_pair_f64
        a = [np.ascontiguousarray(np.asarray(x, float).reshape(-1)) for x in (Ra, pa, sa, Rb, pb, sb)]
        return bool(f(ctypes.c_int(ta), self._p(a[0]), self._p(a[1]), self._p(a[2]), ctypes.c_int(tb), self._p(a[3]),
                      self._p(a[4]), self._p(a[5]), ctypes.c_double(margin)))


@pytest.fixture(scope="session")
def scene():
    import gik_b200
    return gik_b200.nextage_scene()


@pytest.fixture(scope="session")
def scene_c(scene):
    return scene.to_c()


@pytest.fixture(scope="session")
def hostsim():
    return HostSim()


def make_poses(n, seed, box="workspace"):
    """Identity-rotation cube placements, [n,12] float64.  `workspace` = SURVEY 8d config-2 box,
    `sampler` = the reference sampler box (path.py:35-37)."""
    rng = np.random.default_rng(seed)
    lo, hi = {"workspace": ([0.20, -0.40, 0.93], [0.60, 0.40, 1.40]),
              "sampler": ([0.33, -0.30, 1.05], [0.40, 0.11, 1.40])}[box]
    P = np.zeros((n, 12))
    P[:, [0, 4, 8]] = 1.0
    P[:, 9:] = rng.uniform(lo, hi, size=(n, 3))
    return P


def rot_rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return (np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]]) @ np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]]) @
            np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]]))
