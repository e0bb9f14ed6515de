import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_pybullet_drones_b200.envs import HoverAviary, MultiHoverAviary
from gym_pybullet_drones_b200.utils.enums import ActionType, Physics
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
def timed(fn, n, warm=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return round(best, 2)
out = {}
E = 4096
envs = [HoverAviary(physics=Physics.DYN, act=ActionType.PID, num_envs=E, autoreset="same_step") for _ in range(8)]
sp = [(torch.tensor([-0.5, -0.5, 0.5], device=dev) + torch.rand((E, 1, 3), device=dev, generator=g)) for _ in range(8)]
for e in envs: e.reset()
k = [0]
def f2():
    i = k[0] % 8; k[0] += 1; envs[i].step(sp[i])
out["pid4096_us"] = timed(f2, 1000)
del envs
E = 32768
envs = [MultiHoverAviary(num_drones=2, physics=Physics.DYN, act=ActionType.RPM, num_envs=E, autoreset="same_step") for _ in range(8)]
acts = [torch.rand((E, 2, 4), device=dev, generator=g) * 2 - 1 for _ in range(8)]
for e in envs: e.reset()
def f3():
    i = k[0] % 8; k[0] += 1; envs[i].step(acts[i])
out["rpm65536_us"] = timed(f3, 1000)
print(os.environ.get("QS_LIBQUADSIM", "default"), os.environ.get("QS_CTA_CAP"), os.environ.get("QS_PDL"), json.dumps(out))
