import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_pybullet_drones_b200.envs import MultiHoverAviary
from gym_pybullet_drones_b200.utils.enums import ActionType, Physics
E=32768
env=MultiHoverAviary(num_drones=2,physics=Physics.DYN,act=ActionType.RPM,num_envs=E,autoreset="same_step",host_copy=False)
env.reset()
rng=np.random.default_rng(0)
acts=[rng.uniform(-1,1,(E,2,4)).astype(np.float32) for _ in range(4)]
for i in range(30): env.step(acts[i%4])
torch.cuda.synchronize(); t=time.perf_counter()
K=300
for i in range(K): env.step(acts[i%4])
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/K
print(json.dumps({"e2e_ms_per_step":dt*1e3,"drone_steps_per_s":E*2/dt}))
