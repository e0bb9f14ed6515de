"""The CUDA kernels' per-drone core (csrc/quad_core.cuh) compiled for the host and checked against the golden
vectors of the unmodified reference -- same source as the GPU build, same float32-storage/float64-register
split, no GPU needed.  The GPU suite (-m gpu) repeats these through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from gym_pybullet_drones_b200 import _native as N
from gym_pybullet_drones_b200.params import AviaryConstants, fill_params, quaternion_from_euler
from gym_pybullet_drones_b200.utils.enums import DroneModel
from qs_testlib import FIELDS, HostSim, RTOL, host_harness, ptr, quat_err, relerr


def check_fields(out, g, key, t, tol=RTOL, idx=slice(None)):
    for f in FIELDS:
        ref = g[key + "_" + f][t][idx]
        e = quat_err(out[f], ref) if f == "quat" else relerr(out[f], ref)
        assert e <= tol, (key, f, t, e)


@pytest.mark.parametrize("cf,steps", [(240, 1000), (30, 125)])
@pytest.mark.parametrize("stream", ["zeros", "rand", "sine", "const"])
def test_hover_rpm_1000_physics_steps(golden, cf, steps, stream):
    """BASELINE config 1: HoverAviary, 1 drone, DYN, 240 Hz, 1000 physics steps, <= 1e-5 rel on every state element."""
    g = golden("hover_rpm_1000")
    key = "cf%d_%s" % (cf, stream)
    c = AviaryConstants(DroneModel.CF2X, 240, cf)
    sim = HostSim(fill_params(c), 1, N.ACT_RPM, 4, c.PYB_STEPS_PER_CTRL)
    sim.set_state(c.default_init_xyzs(1), quaternion_from_euler(np.zeros((1, 3))), np.zeros((1, 3)), np.zeros((1, 3)))
    acts = g[key + "_actions"]
    tol = RTOL if stream != "const" else 1e-4      # constant torque: |w| reaches 40 rad/s, divergent trajectory
    for t in range(min(steps, acts.shape[0])):
        out = sim.tick(acts[t])
        check_fields(out, g, key, t, tol)


def test_hover_rpm_long_horizon_report(golden):
    """Informational bound for 1000 control ticks at 30 Hz (8000 physics steps, 33 s of open-loop tumbling)."""
    g = golden("hover_rpm_1000")
    c = AviaryConstants(DroneModel.CF2X, 240, 30)
    worst = 0.0
    for stream in ("rand", "sine"):
        key = "cf30_" + stream
        sim = HostSim(fill_params(c), 1, N.ACT_RPM, 4, 8)
        sim.set_state(c.default_init_xyzs(1), quaternion_from_euler(np.zeros((1, 3))), np.zeros((1, 3)), np.zeros((1, 3)))
        acts = g[key + "_actions"]
        for t in range(acts.shape[0]):
            out = sim.tick(acts[t])
            for f in FIELDS:
                ref = g[key + "_" + f][t]
                worst = max(worst, quat_err(out[f], ref) if f == "quat" else relerr(out[f], ref))
    assert worst < 2e-4, worst


@pytest.mark.parametrize("model", [DroneModel.CF2P, DroneModel.RACE])
def test_other_models_raw_rpm(golden, model):
    g = golden("ctrl_models_300")
    key = model.value
    c = AviaryConstants(model, 240, 120)
    sim = HostSim(fill_params(c), 2, N.ACT_RAW_RPM, 4, 2)
    sim.set_state(c.default_init_xyzs(2), quaternion_from_euler(np.zeros((2, 3))), np.zeros((2, 3)), np.zeros((2, 3)))
    acts = g[key + "_actions"]
    for t in range(acts.shape[0]):
        # raw RPM enters the ABI as float32: give both sides the same float32-representable sequence? The
        # golden used float64 rpm, so allow the 6e-8 input rounding on top of the state tolerance.
        out = sim.tick(acts[t])
        obs = g[key + "_obs"][t]
        assert relerr(out["pos"], obs[:, 0:3]) < 3e-5 and quat_err(out["quat"], obs[:, 3:7]) < 3e-5, t
        assert relerr(out["vel"], obs[:, 10:13]) < 3e-5 and relerr(out["ang_v"], obs[:, 13:16]) < 1e-4, t
        assert relerr(out["rpm"], obs[:, 16:20]) < 1e-7


@pytest.mark.parametrize("model", [DroneModel.CF2X, DroneModel.CF2P])
def test_pid_known_answers(golden, model):
    g = golden("pid_kat")
    k = model.value + "_"
    L = host_harness()
    P = fill_params(AviaryConstants(model), pid_model=model)
    n = g[k + "pos"].shape[0]
    st = np.zeros((n, 9))
    for call in range(3):
        for i in range(n):
            rpm, pe, ye = np.zeros(4), np.zeros(3), C.c_double()
            pos = g[k + "pos"][i] + 0.01 * call
            L.hh_pid(C.addressof(P), ptr(st[i]), 1 / 48, ptr(pos), ptr(g[k + "quat"][i].copy()), ptr(g[k + "vel"][i].copy()),
                     ptr(g[k + "target_pos"][i].copy()), float(g[k + "target_rpy"][i, 2]), ptr(g[k + "target_vel"][i].copy()),
                     ptr(g[k + "target_rpy_rates"][i].copy()), ptr(rpm), ptr(pe), C.addressof(ye))
            assert relerr(rpm, g[k + "rpm"][call][i]) < 1e-9, (call, i)
            assert relerr(pe, g[k + "pos_e"][call][i]) < 1e-12 and abs(ye.value - g[k + "yaw_e"][call][i]) < 1e-10
        assert relerr(st[:, 0:3], g[k + "integral_pos_e"][call]) < 1e-12
        assert relerr(st[:, 3:6], g[k + "last_rpy"][call]) < 1e-12
        assert relerr(st[:, 6:9], g[k + "integral_rpy_e"][call]) < 1e-10


ACTS = {"pid": (N.ACT_PID, 3), "vel": (N.ACT_VEL, 4), "one_d_pid": (N.ACT_ONE_D_PID, 1)}


@pytest.mark.parametrize("key,nd,act", [("hover_d1_pid", 1, "pid"), ("hover_d1_vel", 1, "vel"), ("hover_d1_one_d_pid", 1, "one_d_pid"), ("multi_d2_pid", 2, "pid")])
def test_rl_pid_actions_teacher_forced(golden, key, nd, act):
    """240/30 Hz RL envs with the embedded PID: each tick starts from the reference's previous state (the free-running
    30 Hz loop is chaotic in the reference itself, see tests/test_oracle_golden.py)."""
    g = golden("rl_pid_cf30")
    c = AviaryConstants(DroneModel.CF2X, 240, 30)
    at, A = ACTS[act]
    sim = HostSim(fill_params(c), nd, at, A, 8, pid=True)
    acts = g[key + "_actions"]
    for t in range(acts.shape[0]):
        if t == 0:
            sim.set_state(c.default_init_xyzs(nd), quaternion_from_euler(np.zeros((nd, 3))), np.zeros((nd, 3)), np.zeros((nd, 3)))
        else:
            sim.set_state(g[key + "_pos"][t - 1], g[key + "_quat"][t - 1], g[key + "_vel"][t - 1], g[key + "_rpy_rates"][t - 1])
            sim.pid[0:3] = g[key + "_pid_integral_pos_e"][t - 1].T
            sim.pid[3:6] = g[key + "_pid_last_rpy"][t - 1].T
            sim.pid[6:9] = g[key + "_pid_integral_rpy_e"][t - 1].T
        out = sim.tick(acts[t])
        check_fields(out, g, key, t, 2e-5)


@pytest.mark.parametrize("key,nd,act", [("hover_d1_pid", 1, "pid"), ("hover_d1_vel", 1, "vel"), ("hover_d1_one_d_pid", 1, "one_d_pid"), ("multi_d2_pid", 2, "pid")])
def test_rl_pid_actions_trajectory_120hz(golden, key, nd, act):
    """240/120 Hz: contractive closed loop, whole 480-tick trajectories compared."""
    g = golden("rl_pid_cf120")
    c = AviaryConstants(DroneModel.CF2X, 240, 120)
    at, A = ACTS[act]
    sim = HostSim(fill_params(c), nd, at, A, 2, pid=True)
    sim.set_state(c.default_init_xyzs(nd), quaternion_from_euler(np.zeros((nd, 3))), np.zeros((nd, 3)), np.zeros((nd, 3)))
    acts = g[key + "_actions"]
    for t in range(acts.shape[0]):
        out = sim.tick(acts[t])
        check_fields(out, g, key, t, 5e-5)


def test_half_angle_series_matches_libm():
    L = host_harness()
    rng = np.random.default_rng(0)
    for n in np.concatenate([rng.uniform(0, 500, 200), [1e-7, 1e-3, 239.9, 240.1, 1e4]]):
        c, s = C.c_double(), C.c_double()
        L.hh_half_angle(float(n * n), 1 / 240, C.byref(c), C.byref(s))
        th = n / 480
        assert abs(c.value - np.cos(th)) < 3e-16 and abs(s.value - np.sin(th) / n) < 3e-19 + 1e-15 * abs(np.sin(th) / n)
