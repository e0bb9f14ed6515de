"""Throughput of the other BASELINE.json configs (not the headline bench line): writes gpurun_out/configs_r02.json.
    config 2: 4096 x HoverAviary with the embedded DSLPIDControl (act=PID), + stand-alone qs_pid_control
    config 3b: learn.py's ONE_D_RPM variant of the headline workload (A=1, obs 27)
    config 4: one 16384-drone formation with ground effect + downwash (pairwise kernel + one launch per substep)
    S=1:      65536 drones at 240 Hz control (B=120 -> obs 492 floats, 4006 algorithmic bytes per drone-step)
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_pybullet_drones_b200.control import DSLPIDControl
from gym_pybullet_drones_b200.envs import CtrlAviary, HoverAviary, MultiHoverAviary
from gym_pybullet_drones_b200.utils.enums import ActionType, DroneModel, Physics

PEAK = 6572.5
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, n, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {}
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)

# config 2: several independent 4096-env batches rotated (a single 4096 batch is 2.4 MB: launch-latency bound, L2 resident)
E = 4096
envs = [HoverAviary(physics=Physics.DYN, act=ActionType.PID, num_envs=E, autoreset="same_step") for _ in range(8)]
sp = [(torch.tensor([-0.5, -0.5, 0.5], device=dev) + torch.rand((E, 1, 3), device=dev, generator=g)) for _ in range(8)]
for e in envs:
    e.reset()
k = [0]
def f2():
    i = k[0] % 8; k[0] += 1
    envs[i].step(sp[i])
ms = timed(f2, 2000, 50)
out["config2_hover_pid_4096"] = {"drones": E, "S": 8, "ms_per_step": ms, "drone_steps_per_s": E / (ms * 1e-3), "alg_bytes": 598,
                                  "hbm_frac": 598 * E / (ms * 1e-3) / 1e9 / PEAK, "note": "4096 drones = 2.4 MB per launch: launch-latency bound"}
for n in (4096, 1048576):
    ctrl = DSLPIDControl(DroneModel.CF2X, num_drones=n)
    pos = torch.rand((n, 3), device=dev, generator=g); vel = torch.rand((n, 3), device=dev, generator=g) - 0.5
    q = torch.randn((n, 4), device=dev, generator=g); q = q / q.norm(dim=1, keepdim=True)
    tp = torch.rand((n, 3), device=dev, generator=g)
    ms = timed(lambda: ctrl.computeControl(1 / 48, pos, q, vel, None, tp), 200, 10)
    out["qs_pid_control_%d" % n] = {"n": n, "ms_per_call": ms, "calls_per_s": n / (ms * 1e-3), "alg_bytes": 192, "hbm_frac": 192 * n / (ms * 1e-3) / 1e9 / PEAK}
    # the same controller reading the float64 state planes of an env (qs_pid_control_state): coalesced 32-byte loads, float64 RPMs out
    cenv = CtrlAviary(num_drones=1, physics=Physics.DYN, num_envs=n)
    cenv.reset()
    tp64 = tp.double()
    ms = timed(lambda: ctrl.computeControlFromEnv(cenv, tp64, control_timestep=1 / 48), 200, 10)
    actual = 96 + 72 + 72 + 24 + 32 + 16          # planes r, pid state r/w (float64), target r, rpm w (float64), pos_e/yaw_e w
    out["qs_pid_control_state_%d" % n] = {"n": n, "ms_per_call": ms, "calls_per_s": n / (ms * 1e-3), "alg_bytes": 192, "actual_bytes": actual,
                                          "hbm_frac": 192 * n / (ms * 1e-3) / 1e9 / PEAK, "hbm_frac_actual_bytes": actual * n / (ms * 1e-3) / 1e9 / PEAK}
    del cenv

# config 3b: ONE_D_RPM
E, D = 32768, 2
envs = [MultiHoverAviary(num_drones=D, physics=Physics.DYN, act=ActionType.ONE_D_RPM, num_envs=E, autoreset="same_step") for _ in range(8)]
acts = [torch.rand((E, D, 1), device=dev, generator=g) * 2 - 1 for _ in range(8)]
for e in envs:
    e.reset()
k = [0]
def f3():
    i = k[0] % 8; k[0] += 1
    envs[i].step(acts[i])
ms = timed(f3, 2000, 50)
out["config3_multihover_one_d_rpm_65536"] = {"drones": E * D, "S": 8, "ms_per_step": ms, "drone_steps_per_s": E * D / (ms * 1e-3), "alg_bytes": 286,
                                             "hbm_frac": 286 * E * D / (ms * 1e-3) / 1e9 / PEAK}
del envs

# config 3 (ii): SB3-MlpPolicy-shaped torch network + env entirely on the device, captured in a CUDA graph
E, D = 32768, 2
env = MultiHoverAviary(num_drones=D, physics=Physics.DYN, act=ActionType.RPM, num_envs=E, autoreset="same_step")
torch.manual_seed(0)
pi = torch.nn.Sequential(torch.nn.Linear(D * 72, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(), torch.nn.Linear(64, D * 4)).cuda()
obs, _ = env.reset()
act = torch.zeros((E, D, 4), device=dev)
def pol_step(o):
    with torch.no_grad():
        mean = pi(o.reshape(E, -1))
        act.copy_((mean + 0.3 * torch.randn_like(mean)).clamp(-1, 1).view(E, D, 4))
    return env.step(act)[0]
for _ in range(4):
    obs = pol_step(obs)
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
views = [env._obs_view[0], env._obs_view[1]]
with torch.cuda.graph(gr):
    pol_step(views[env._cur]); pol_step(views[env._cur])
ms = timed(gr.replay, 200, 10) / 2
out["config3_policy_plus_env_cuda_graph_65536"] = {"drones": E * D, "S": 8, "ms_per_step": ms, "drone_steps_per_s": E * D / (ms * 1e-3),
                                                   "note": "fp32 MLP 144-64-64-8 (tanh) with Gaussian exploration noise + fused env step, 2 steps per graph replay"}
del env, gr

# S=1: 240 Hz control, B=120
E, D = 32768, 2
env = MultiHoverAviary(num_drones=D, physics=Physics.DYN, act=ActionType.RPM, pyb_freq=240, ctrl_freq=240, num_envs=E, autoreset="same_step")
a = torch.rand((E, D, 4), device=dev, generator=g) * 2 - 1
env.reset()
ms = timed(lambda: env.step(a), 500, 20)
out["rpm_240hz_S1_65536"] = {"drones": E * D, "S": 1, "obs_dim": 492, "ms_per_step": ms, "drone_steps_per_s": E * D / (ms * 1e-3), "alg_bytes": 4006,
                             "hbm_frac": 4006 * E * D / (ms * 1e-3) / 1e9 / PEAK, "note": "working set 258 MB > L2; unstaged writer (span too large for shared memory)"}
del env

# config 4: one aviary of 16384 drones, 128x128 grid 0.15 m pitch, z = 0.1 + 0.05*(i mod 16), GND|DW
Dn = 16384
i = np.arange(Dn)
xyz = np.stack([0.15 * (i % 128), 0.15 * (i // 128), 0.1 + 0.05 * (i % 16)], axis=1)
env = CtrlAviary(num_drones=Dn, initial_xyzs=xyz, physics=Physics.PYB_GND_DRAG_DW, pyb_freq=240, ctrl_freq=240, num_envs=1)
env.reset()
rpm = torch.full((1, Dn, 4), float(env.HOVER_RPM), device=dev)
ms = timed(lambda: env.step(rpm), 50, 5)
import ctypes as C
from gym_pybullet_drones_b200 import _native as N
import os
from gym_pybullet_drones_b200.formation import morton_order


def dw_ms(e, cull, boxed=True):
    os.environ["QS_DW_CULL"] = "1" if cull else "0"
    f = torch.zeros(e._D, device=dev)
    ws = torch.zeros(((e._D + 31) // 32, 8), device=dev)
    sp = torch.cuda.current_stream().cuda_stream
    if boxed:
        t = timed(lambda: N.lib().qs_downwash_boxed(C.byref(e._P), C.byref(e._st), 1, e._D, ws.data_ptr(), f.data_ptr(), sp), 50, 5)
    else:
        t = timed(lambda: N.lib().qs_downwash(C.byref(e._P), C.byref(e._st), 1, e._D, f.data_ptr(), sp), 50, 5)
    os.environ["QS_DW_CULL"] = "1"
    return t


env.reset()                                   # the force kernel is timed on the formation's reset geometry
ms_all, ms_cull = dw_ms(env, False), dw_ms(env, True)
ms_all_tiled, ms_cull_tiled = dw_ms(env, False, boxed=False), dw_ms(env, True, boxed=False)
env_m = CtrlAviary(num_drones=Dn, initial_xyzs=xyz[morton_order(xyz[:, :2])], physics=Physics.PYB_GND_DRAG_DW, pyb_freq=240, ctrl_freq=240, num_envs=1)
env_m.reset()
ms_morton, ms_morton_tiled = dw_ms(env_m, True), dw_ms(env_m, True, boxed=False)
ms_step_morton = timed(lambda: (env_m.reset(), env_m.step(rpm)), 20, 3)
out["config4_formation_16384_gnd_drag_dw"] = {"drones": Dn, "S": 1, "ms_per_step": ms, "drone_steps_per_s": Dn / (ms * 1e-3),
                                              "downwash_kernel_ms": ms_cull, "pairs_per_s": Dn * Dn / (ms_cull * 1e-3),
                                              "downwash_ms": {"all_pairs": ms_all, "culled_row_major": ms_cull, "culled_morton_order": ms_morton,
                                                              "tiled_kernel_all_pairs": ms_all_tiled, "tiled_kernel_culled_row_major": ms_cull_tiled,
                                                              "tiled_kernel_culled_morton_order": ms_morton_tiled},
                                              "reset_plus_step_ms_morton_order": ms_step_morton,
                                              "pairs_per_s_all_pairs": Dn * Dn / (ms_all * 1e-3),
                                              "note": "pairwise term is FP32 ALU/SFU bound, not HBM bound; qs_downwash_boxed: boxes kernel + box-table kernel; exact culling skips chunks that cannot contribute (pairs/s counts all N^2 pairs); ms_per_step = steps from a drifting formation"}
adj = torch.empty((1, Dn, Dn), dtype=torch.uint8, device=dev)
env.NEIGHBOURHOOD_RADIUS = 1.0
ms_adj = timed(lambda: env.adjacency(adj), 50, 5)
out["adjacency_16384"] = {"drones": Dn, "S": 1, "ms_per_step": ms_adj, "drone_steps_per_s": Dn / (ms_adj * 1e-3), "alg_bytes": Dn,
                          "hbm_frac": Dn * Dn / (ms_adj * 1e-3) / 1e9 / PEAK, "note": "qs_adjacency: N^2 bytes written per query (268 MB)"}
del env_m, adj
# a formation 4x larger: 65536 drones, 256x256 grid, Morton order
Db = 65536
ib = np.arange(Db)
xb = np.stack([0.15 * (ib % 256), 0.15 * (ib // 256), 0.1 + 0.05 * (ib % 16)], axis=1)
env_b = CtrlAviary(num_drones=Db, initial_xyzs=xb[morton_order(xb[:, :2])], physics=Physics.PYB_GND_DRAG_DW, pyb_freq=240, ctrl_freq=240, num_envs=1)
env_b.reset()
ms_b = dw_ms(env_b, True)
out["formation_65536_downwash_culled_morton"] = {"drones": Db, "S": 1, "ms_per_step": ms_b, "drone_steps_per_s": Db / (ms_b * 1e-3),
                                                  "pairs_per_s": Db * Db / (ms_b * 1e-3), "note": "downwash kernel only; 4.3e9 pairs, culled"}
json.dump(out, open("gpurun_out/configs_r02.json", "w"), indent=1)
print(json.dumps(out, indent=1))
