import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_pybullet_drones_b200.envs import MultiHoverAviary
from gym_pybullet_drones_b200.utils.enums import ActionType, Physics
dev = torch.device("cuda:0")
n = 65536 * 72
d = torch.zeros(n, device=dev)
h = torch.zeros(n).pin_memory()
for name, f in (("D2H 18.9MB pinned", lambda: h.copy_(d, non_blocking=True)), ("H2D 18.9MB pinned", lambda: d.copy_(h, non_blocking=True))):
    for _ in range(3): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): f()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    print("%s: %.3f ms  %.1f GB/s" % (name, dt * 1e3, n * 4 / dt / 1e9))
hp = torch.zeros(n)   # pageable
t0 = time.perf_counter()
for _ in range(5): hp.copy_(d)
print("D2H pageable: %.3f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
E, D = 32768, 2
env = MultiHoverAviary(num_drones=D, physics=Physics.DYN, act=ActionType.RPM, num_envs=E, autoreset="same_step", host_copy=False)
env.reset()
a = np.random.default_rng(0).uniform(-1, 1, (E, D, 4)).astype(np.float32)
for _ in range(5): env.step(a)
t0 = time.perf_counter()
for _ in range(50): env.step(a)
print("env.step(numpy): %.3f ms" % ((time.perf_counter() - t0) / 50 * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(50): env.step(a)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
print("cpu count", os.cpu_count())
