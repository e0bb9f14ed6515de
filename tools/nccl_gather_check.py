"""torchrun --nproc-per-node G tools/nccl_gather_check.py : every rank steps its slice of the aviaries (no collective on
the step path), then NCCL all-gathers observations/rewards; rank 0 checks the result is bit-identical to one GPU stepping
all aviaries, and reports the cost of the optional all-gather."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from gym_pybullet_drones_b200.envs import MultiHoverAviary
from gym_pybullet_drones_b200.sharding import all_gather_envs, shard_envs
from gym_pybullet_drones_b200.utils.enums import ActionType, Physics

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)
dist.init_process_group("nccl", device_id=dev)
dist.all_reduce(torch.zeros(1, device=dev)); torch.cuda.synchronize()
sys.stdout.flush(); os.dup2(saved, 1)
E, D, T = 32768 * world, 2, 20
sh = shard_envs(E)
g = torch.Generator(device="cpu").manual_seed(0)
acts = (torch.rand((T, E, D, 4), generator=g) * 2 - 1)
env = MultiHoverAviary(num_drones=D, physics=Physics.DYN, act=ActionType.RPM, num_envs=sh.count, device=dev, autoreset="same_step")
env.reset()
mine = acts[:, sh.start:sh.stop].to(dev)
for t in range(T):
    obs, rew, term, trunc, _ = env.step(mine[t])
g_obs = all_gather_envs(obs, sh); g_rew = all_gather_envs(rew, sh); g_done = all_gather_envs((term | trunc).to(torch.uint8), sh)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier(); e0.record()
for _ in range(20):
    g_obs = all_gather_envs(obs, sh)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
if rank == 0:
    ref = MultiHoverAviary(num_drones=D, physics=Physics.DYN, act=ActionType.RPM, num_envs=E, device=dev, autoreset="same_step")
    ref.reset()
    full = acts.to(dev)
    for t in range(T):
        o, r, te, tr, _ = ref.step(full[t])
    ok = bool(torch.equal(o, g_obs)) and bool(torch.equal(r, g_rew)) and bool(torch.equal((te | tr).to(torch.uint8), g_done))
    print(json.dumps({"world": world, "sharded_equals_single_gpu": ok, "all_gather_obs_ms": ms,
                      "all_gather_GBps_per_gpu": g_obs.numel() * 4 / (ms * 1e-3) / 1e9}))
dist.destroy_process_group()
