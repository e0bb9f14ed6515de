import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_pybullet_drones_b200.envs import MultiHoverAviary
from gym_pybullet_drones_b200.utils.enums import ActionType, Physics
E, D = 4096, 2
env = MultiHoverAviary(num_drones=D, physics=Physics.DYN, act=ActionType.RPM, num_envs=E, autoreset="same_step", host_copy=False)
env.reset()
rng = np.random.default_rng(0)
for t in range(60):
    a = rng.uniform(-1, 1, (E, D, 4)).astype(np.float32)
    o, r, te, tr, info = env.step(a)
    nd = int((te | tr).sum())
    print(t, "done", nd, "n_final", info.get("final_obs", np.zeros((0,))).shape[0], "dev done", int(env._done.sum().item()), "nfinal_dev", int(env._nfinal_dev.item()), env._h_nfinal[0].item(), env._h_nfinal[1].item()) if t % 6 == 0 or nd else None
