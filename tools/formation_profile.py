"""A few launches of the formation kernels at config-4 size for `ncu --set full -k regex:downwash_boxed|adjacency|dw_boxes`."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_pybullet_drones_b200 import _native as N  # noqa: E402
from gym_pybullet_drones_b200.envs import CtrlAviary  # noqa: E402
from gym_pybullet_drones_b200.formation import morton_order  # noqa: E402
from gym_pybullet_drones_b200.utils.enums import Physics  # noqa: E402

Dn = 16384
i = np.arange(Dn)
xyz = np.stack([0.15 * (i % 128), 0.15 * (i // 128), 0.1 + 0.05 * (i % 16)], axis=1)
env = CtrlAviary(num_drones=Dn, initial_xyzs=xyz[morton_order(xyz[:, :2])], neighbourhood_radius=1.0, physics=Physics.PYB_GND_DRAG_DW,
                 pyb_freq=240, ctrl_freq=240, num_envs=1)
env.reset()
fz = torch.zeros(Dn, device="cuda")
ws = torch.zeros(((Dn + 31) // 32, 8), device="cuda")
adj = torch.empty((1, Dn, Dn), dtype=torch.uint8, device="cuda")
sp = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    N.check(N.lib().qs_downwash_boxed(C.byref(env._P), C.byref(env._st), 1, Dn, ws.data_ptr(), fz.data_ptr(), sp), "boxed")
    env.adjacency(adj)
torch.cuda.synchronize()
print("ok", float(fz.abs().max()), int(adj.sum()))
