import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_pybullet_drones_b200.envs import MultiHoverAviary
from gym_pybullet_drones_b200.utils.enums import ActionType, Physics
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
env = MultiHoverAviary(num_drones=2, physics=Physics.DYN, act=ActionType.RPM, num_envs=n // 2, autoreset="same_step")
a = torch.rand((n // 2, 2, 4), device="cuda") * 2 - 1
env.reset()
for _ in range(12): env.step(a)
torch.cuda.synchronize()
