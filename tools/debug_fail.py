import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from gym_pybullet_drones_b200.envs import CtrlAviary, MultiHoverAviary
from gym_pybullet_drones_b200.utils.enums import ActionType, Physics
from oracle import dyn_oracle as O
from qs_testlib import relerr, quat_err
from test_gpu_parity import state_of

def cmp(env, ora, tag, t):
    st = state_of(env)
    out = []
    for f in ("pos", "quat", "rpy", "vel", "ang_v", "rpy_rates"):
        ref = getattr(ora, f)
        a = st[f]
        if f == "quat":
            s = np.sign(np.sum(a * ref, -1, keepdims=True)); a = a * s
        e = np.abs(a - ref) / np.maximum(np.abs(ref), 1)
        idx = np.unravel_index(np.argmax(e), e.shape)
        out.append("%s %.1e@%s" % (f, e.max(), idx))
    print(tag, t, " ".join(out), flush=True)

# case 1: config4 big formation
D = 1024
i = np.arange(D)
xyz = np.stack([0.25 * (i % 32), 0.25 * (i // 32), 0.5 + 0.5 * (i % 4)], axis=1)
for phys, eff in (("PYB_DW", 4), ("PYB_GND", 1), ("PYB_DRAG", 2), ("PYB_GND_DRAG_DW", 7), ("DYN", 0)):
    env = CtrlAviary(num_drones=D, initial_xyzs=xyz, physics=Physics[phys], pyb_freq=240, ctrl_freq=48, num_envs=1)
    ora = O.OracleAviary("ctrl", 1, D, ctrl_freq=48, initial_xyzs=xyz, effects=eff)
    rng = np.random.default_rng(9)
    env.reset(); ora.reset()
    for t in range(3):
        a = (ora.P.HOVER_RPM * (1 + 0.05 * rng.uniform(-1, 1, (1, D, 4)))).astype(np.float32)
        obs, *_ = env.step(torch.from_numpy(a).cuda()); ora.step(a)
        cmp(env, ora, "big " + phys, t)

# case 2: small-D combined effects
E, Dn, T = 64, 4, 60
xyz = np.array([[0.0, 0.0, 0.06], [0.05, 0.02, 0.4], [0.1, -0.03, 0.75], [-0.05, 0.05, 1.1]])
rng = np.random.default_rng(5)
acts = (0.3 * rng.uniform(-1, 1, (T, E, Dn, 4))).astype(np.float32)
env = MultiHoverAviary(num_drones=Dn, initial_xyzs=xyz, physics=Physics.PYB_GND_DRAG_DW, act=ActionType.RPM, num_envs=E)
ora = O.OracleAviary("multihover", E, Dn, act="rpm", initial_xyzs=xyz, effects=7)
env.reset(); ora.reset()
for t in range(T):
    obs, rew, term, trunc, _ = env.step(torch.from_numpy(acts[t]).cuda())
    o_obs, o_rew, o_term, o_trunc = ora.step(acts[t])
    if t % 6 == 0 or t > 50: cmp(env, ora, "small7", t)
    if not np.array_equal(trunc.cpu().numpy(), o_trunc): print("trunc mismatch", t, np.nonzero(trunc.cpu().numpy() != o_trunc))

# case 3: config3
E, Dn, T = 2048, 2, 125
rng = np.random.default_rng(123)
acts = rng.uniform(-1, 1, (T, E, Dn, 4)).astype(np.float32)
env = MultiHoverAviary(num_drones=Dn, physics=Physics.DYN, act=ActionType.RPM, num_envs=E)
ora = O.OracleAviary("multihover", E, Dn, act="rpm")
env.reset(); ora.reset()
for t in range(T):
    obs, rew, term, trunc, _ = env.step(torch.from_numpy(acts[t]).cuda())
    o_obs, o_rew, o_term, o_trunc = ora.step(acts[t])
    if t % 10 == 0 or t > 118: cmp(env, ora, "cfg3", t)
    if not np.array_equal(trunc.cpu().numpy(), o_trunc): print("trunc mismatch", t, np.nonzero(trunc.cpu().numpy() != o_trunc))
    re = np.abs(rew.cpu().numpy() - o_rew).max()
    if re > 1e-5: print("rew err", t, re)
