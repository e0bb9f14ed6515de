import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_pybullet_drones_b200.envs import MultiHoverAviary
from gym_pybullet_drones_b200.utils.enums import ActionType, Physics
E, D = 32768, 2
env = MultiHoverAviary(num_drones=D, physics=Physics.DYN, act=ActionType.RPM, num_envs=E, autoreset="same_step", host_copy=False)
env.reset()
a = np.random.default_rng(0).uniform(-1, 1, (E, D, 4)).astype(np.float32)
for _ in range(10): env.step(a)
t0 = time.perf_counter()
for _ in range(100): env.step(a)
print("env.step(numpy): %.3f ms" % ((time.perf_counter() - t0) / 100 * 1e3))
# stage timings (each stage followed by a sync)
def T(f, n=50):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("host memcpy action -> pinned: %.3f ms" % T(lambda: env._h_action.numpy().__setitem__(Ellipsis, a.reshape(-1, 4))))
print("H2D action: %.3f ms" % T(lambda: env._action_dev.copy_(env._h_action, non_blocking=True)))
print("launch (generic path): %.3f ms" % T(lambda: env._launch(env._action_dev)))
print("D2H obs: %.3f ms" % T(lambda: env._h_obs[0].copy_(env._obs_buf[0], non_blocking=True)))
print("D2H 3 small: %.3f ms" % T(lambda: (env._h_reward[0].copy_(env._reward, non_blocking=True), env._h_term[0].copy_(env._terminated, non_blocking=True), env._h_trunc[0].copy_(env._truncated, non_blocking=True))))
idx = np.arange(0, E, 100)
def fin():
    idx_dev = torch.from_numpy(idx).to(env.device, non_blocking=True)
    return torch.index_select(env._final_view, 0, idx_dev).cpu().numpy()
print("final rows (328 envs): %.3f ms" % T(fin))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(50): env.step(a)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(10)
