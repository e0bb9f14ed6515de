"""Shared helpers for the parity tests (test infrastructure)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = ("pos", "quat", "rpy", "vel", "ang_v", "rpy_rates")

# north-star tolerance: |a-b| <= 1e-5 * max(|b|, 1) element-wise on the kinematic state
RTOL = 1e-5
# what the float64 state planes actually deliver on non-chaotic trajectories (the kernels differ from the float64
# reference only in operation order, series vs libm in _integrateQ, and 1/|q| renormalisation): used where the
# observation's float32 cast is not in the way
TIGHT = 1e-9


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0))) if a.size else 0.0


def quat_err(a, b):
    """sign-insensitive quaternion comparison (q and -q are the same rotation)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    s = np.sign(np.sum(a * b, axis=-1, keepdims=True))
    s[s == 0] = 1
    return relerr(a * s, b)


def pack_planes(pos, quat, vel, w):
    """float64 [n,3],[n,4],[n,3],[n,3] -> float64 planes [13 n] in the layout of include/quadsim.h."""
    n = pos.shape[0]
    pl = np.zeros((13 * n,), np.float64)
    p = pl[:12 * n].reshape(3, n, 4)
    p[0, :, 0:3], p[0, :, 3] = pos, w[:, 0]
    p[1] = quat
    p[2, :, 0:3], p[2, :, 3] = vel, w[:, 1]
    pl[12 * n:] = w[:, 2]
    return pl


def unpack_planes(pl):
    pl = np.asarray(pl, np.float64).reshape(-1)
    n = pl.size // 13
    p = pl[:12 * n].reshape(3, n, 4)
    w = np.stack([p[0, :, 3], p[2, :, 3], pl[12 * n:]], axis=1)
    return p[0, :, 0:3], p[1], p[2, :, 0:3], w


_HH = None


def host_harness():
    """Builds (g++) and loads tests/host_harness: the CUDA kernels' per-drone core compiled for the host."""
    global _HH
    if _HH is not None:
        return _HH
    src = os.path.join(ROOT, "tests", "host_harness", "core_host.cpp")
    out = os.path.join(ROOT, "tests", "host_harness", "libcore_host.so")
    deps = [src, os.path.join(ROOT, "gym_pybullet_drones_b200", "csrc", "quad_core.cuh"), os.path.join(ROOT, "include", "quadsim.h")]
    if not os.path.isfile(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", "-o", out, src], check=True)
    L = C.CDLL(out)
    L.hh_downwash_pair.restype = C.c_double
    L.hh_downwash_pair.argtypes = [C.c_void_p, C.c_double, C.c_double]
    L.hh_half_angle.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.hh_tick.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hh_pid.argtypes = [C.c_void_p] + [C.c_void_p, C.c_double] + [C.c_void_p] * 4 + [C.c_double] + [C.c_void_p] * 5
    _HH = L
    return L


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class HostSim:
    """n independent drones stepped by the host build of the kernel core (float64 planes like the CUDA kernels)."""

    def __init__(self, params, n, act_type, A, substeps, effects=0, pid=False):
        self.L = host_harness()
        self.P, self.n, self.act_type, self.A, self.S, self.effects = params, n, act_type, A, substeps, effects
        self.planes = np.zeros((13 * n,), np.float64)
        self.last_rpm = np.zeros((n, 4), np.float64)
        self.pid = np.zeros((9, n), np.float64) if pid else None
        self.rec = np.zeros((n, 23), np.float64)

    def set_state(self, pos, quat, vel, w):
        self.planes[...] = pack_planes(np.asarray(pos, np.float64).reshape(self.n, 3), np.asarray(quat, np.float64).reshape(self.n, 4),
                                       np.asarray(vel, np.float64).reshape(self.n, 3), np.asarray(w, np.float64).reshape(self.n, 3))

    def tick(self, action):
        a = np.ascontiguousarray(np.asarray(action, np.float32).reshape(self.n, self.A))
        self.L.hh_tick(C.addressof(self.P), ptr(self.planes), self.n, ptr(a), self.A, self.act_type, self.S, self.effects,
                       ptr(self.last_rpm), ptr(self.pid), ptr(self.rec))
        r = self.rec
        return dict(pos=r[:, 0:3], quat=r[:, 3:7], rpy=r[:, 7:10], vel=r[:, 10:13], ang_v=r[:, 13:16], rpy_rates=r[:, 16:19], rpm=r[:, 19:23])
