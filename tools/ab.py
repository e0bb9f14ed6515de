import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_pybullet_drones_b200.envs import MultiHoverAviary
from gym_pybullet_drones_b200.utils.enums import ActionType, Physics
dev = torch.device("cuda:0"); D = 2
g = torch.Generator(device=dev).manual_seed(0)
out = {}
for n in (65536, 262144, 1048576):
    envs = [MultiHoverAviary(num_drones=D, physics=Physics.DYN, act=ActionType.RPM, num_envs=n // D, autoreset="same_step") for _ in range(8 if n == 65536 else 1)]
    acts = [torch.rand((n // D, D, 4), device=dev, generator=g) * 2 - 1 for _ in envs]
    for e in envs: e.reset()
    R = len(envs)
    for k in range(40): envs[k % R].step(acts[k % R])
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 400 if n == 65536 else 100
        e0.record()
        for k in range(K): envs[k % R].step(acts[k % R])
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / K * 1e3)
    out[n] = round(best, 2)
    if n == 65536:
        ro = None
        for k in range(3): ro = envs[0].rollout(num_steps=32, seed=1, out=ro)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(20): ro = envs[k % R].rollout(num_steps=32, seed=1, out=ro)
        e1.record(); torch.cuda.synchronize()
        out["rollout_us_per_tick"] = round(e0.elapsed_time(e1) / 640 * 1e3, 2)
        del ro
    del envs, acts
print(os.environ.get("QS_LIBQUADSIM"), json.dumps(out))
