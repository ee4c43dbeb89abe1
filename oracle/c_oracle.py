"""ctypes loader for the C oracle (oracle/grasp_ik_oracle.c) -- TEST INFRASTRUCTURE ONLY.
Row-major numpy in/out, fp64.  The table argument is any ctypes struct laid out like gik_table_t."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _cpu_tag():
    """The oracle is compiled with -march=native, so the artefact is keyed by the host CPU's model and ISA flags:
    a library built in the build container is not reused on a GPU box with a different CPU."""
    import hashlib
    key = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith(("model name", "flags")):
                key += line
            if line.startswith("flags"):
                break
    except OSError:
        pass
    return hashlib.sha1(key.encode()).hexdigest()[:10]


LIB = os.path.join(_HERE, "_build", f"liboracle-{_cpu_tag()}.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "grasp_ik_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", f"OUT={LIB}"], check=True, capture_output=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dt=np.float64):
    return np.ascontiguousarray(a, dtype=dt)


def fk(table_c, q):
    q = _c(q); n = q.shape[0]
    out = np.empty((n, 2, 12))
    lib().orc_fk(ctypes.byref(table_c), ctypes.c_int64(n), _p(q), _p(out))
    return out[:, :, :9].reshape(n, 2, 3, 3), out[:, :, 9:]


def jac(table_c, q):
    q = _c(q); n, nq = q.shape
    out = np.empty((n, 2, 6, nq))
    lib().orc_jac(ctypes.byref(table_c), ctypes.c_int64(n), _p(q), _p(out))
    return out


def solve(table_c, q_init, pose12, eps=1e-3, dt=1e-2, max_iters=1000, threads=0):
    q_init = _c(q_init); pose12 = _c(pose12); n, nq = q_init.shape
    q = np.empty((n, nq)); conv = np.empty(n, np.uint8); it = np.empty(n, np.int32); res = np.empty((n, 2))
    lib().orc_solve(ctypes.byref(table_c), ctypes.c_int64(n), _p(q_init), _p(pose12), ctypes.c_double(eps),
                    ctypes.c_double(dt), ctypes.c_int(max_iters), ctypes.c_int(threads), _p(q), _p(conv), _p(it), _p(res))
    return q, conv.astype(bool), it, res


def interpolate(A12, B12, alpha):
    A12 = _c(A12); B12 = _c(B12); alpha = _c(alpha); n = A12.shape[0]
    out = np.empty((n, 12))
    lib().orc_interpolate(ctypes.c_int64(n), _p(A12), _p(B12), _p(alpha), _p(out))
    return out


def project_edges(table_c, q_start, pose_a, pose_b, num_steps, max_steps, eps=1e-3, dt=1e-2, max_iters=1000, threads=0):
    q_start = _c(q_start); pose_a = _c(pose_a); pose_b = _c(pose_b); ns = _c(num_steps, np.int32)
    n, nq = q_start.shape
    path = np.zeros((n, max_steps, nq)); nv = np.empty(n, np.int32); itt = np.empty(n, np.int32)
    lib().orc_project_edges(ctypes.byref(table_c), ctypes.c_int64(n), ctypes.c_int(max_steps), _p(q_start), _p(pose_a),
                            _p(pose_b), _p(ns), ctypes.c_double(eps), ctypes.c_double(dt), ctypes.c_int(max_iters),
                            ctypes.c_int(threads), _p(path), _p(nv), _p(itt))
    return path, nv, itt


def max_threads():
    return int(lib().orc_max_threads())
