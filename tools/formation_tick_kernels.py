"""Which kernels make up one control tick of a large formation (tools, not product): run under
ncu --metrics gpu__time_duration.sum --clock-control none --csv."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_pybullet_drones_b200.formation import FormationShard, morton_order
from gym_pybullet_drones_b200.utils.enums import Physics
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
side = int(round(n ** 0.5))
i = np.arange(side * side)
xyz = np.stack([0.15 * (i % side), 0.15 * (i // side), 0.1 + 0.05 * (i % 16)], axis=1)
xyz = xyz[morton_order(xyz[:, :2])]
env = FormationShard(xyz, exchange="local", rank=0, world=1, physics=Physics.PYB_GND_DRAG_DW, pyb_freq=240, ctrl_freq=48)
env.reset()
a = torch.full((1, len(xyz), 4), 14468.429183500699, dtype=torch.float32, device=env.device)
torch.cuda.synchronize()
for _ in range(2):
    env.step(a)
torch.cuda.synchronize()
