"""Times qs_rollout with the on-device MlpPolicy at the bench size (tools, not product)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_pybullet_drones_b200.envs import MultiHoverAviary
from gym_pybullet_drones_b200.policy import MlpPolicy
from gym_pybullet_drones_b200.utils.enums import ActionType, Physics
dev = torch.device("cuda:0")
E, D, T = 32768, 2, 16
envs = [MultiHoverAviary(num_drones=D, physics=Physics.DYN, act=ActionType.RPM, num_envs=E, autoreset="same_step") for _ in range(4)]
for e in envs:
    e.reset()
g = torch.Generator(device=dev).manual_seed(0)
noise = torch.randn((T, E, D * 4), device=dev, generator=g)
out = {}
for name, critic in (("env_only", None), ("actor_only", False), ("actor_critic", True)):
    pol = None if critic is None else MlpPolicy.random(D * 72, D * 4, seed=3, critic=critic)
    hold = [None]
    def run(n):
        for k in range(n):
            hold[0] = envs[k % 4].rollout(policy=pol, noise=noise, out=hold[0]) if pol is not None else envs[k % 4].rollout(num_steps=T, seed=1, out=hold[0])
    run(3); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(10); e1.record(); torch.cuda.synchronize()
    out[name] = {"us_per_tick": e0.elapsed_time(e1) / 10 / T * 1e3}
print(json.dumps(out))
