import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_pybullet_drones_b200.envs import MultiHoverAviary
from gym_pybullet_drones_b200.utils.enums import ActionType, Physics
kw = dict(num_drones=32, physics=Physics.DYN, act=ActionType.RPM, num_envs=5, autoreset=None, rpy_f32=True)
e1, e2 = MultiHoverAviary(**kw), MultiHoverAviary(**kw)
e1.reset(); e2.reset()
g = torch.Generator(device="cuda").manual_seed(3)
a = torch.rand((5, 32, 4), device="cuda", generator=g) * 2 - 1
os.environ["QS_FAST"] = "1"; e1.step(a)
os.environ["QS_FAST"] = "0"; e2.step(a)
n = e1._N
d = (e1._planes - e2._planes)
idx = d.nonzero().flatten().cpu().numpy()
print("n diff", len(idx))
for k in idx[:20]:
    if k < 12 * n:
        pl, r = divmod(int(k), 4 * n); dr, c = divmod(r, 4)
        print("plane", pl, "drone", dr, "col", c, float(e1._planes[k]), float(e2._planes[k]), float(d[k]))
    else:
        print("wz drone", int(k) - 12 * n, float(e1._planes[k]), float(e2._planes[k]))
