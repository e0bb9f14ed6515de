"""torchrun --nproc-per-node G tools/gather_multi_gpu.py : the fused observation gather (sharding.ObsGather: every rank's
step kernel stores its rows into the learner's tensor over NVLink and raises a flag) against the baseline (step, then NCCL
all_gather_into_tensor).  Rank 0 = learner: checks both give the same bits and reports the time of a tick as the learner
sees it (step launched -> all rows of all ranks usable on its stream), max over 3 repetitions of the median of 50 ticks."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NCCL_DEBUG", "INFO")
os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT,GRAPH,P2P")
os.environ.setdefault("NCCL_DEBUG_FILE", os.path.join("gpurun_out", "nccl_rank%s.log" % os.environ.get("RANK", "0")))
import torch
import torch.distributed as dist
from gym_pybullet_drones_b200.envs import MultiHoverAviary
from gym_pybullet_drones_b200.sharding import ObsGather, all_gather_envs, shard_envs
from gym_pybullet_drones_b200.utils.enums import ActionType, Physics

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)
dist.init_process_group("nccl", device_id=dev)
dist.all_reduce(torch.zeros(1, device=dev)); torch.cuda.synchronize()
sys.stdout.flush(); os.dup2(saved, 1)
E_per, D = 32768, 2
E = E_per * world
sh = shard_envs(E)
kw = dict(num_drones=D, physics=Physics.DYN, act=ActionType.RPM, device=dev, autoreset="same_step")
env = MultiHoverAviary(num_envs=sh.count, **kw)       # stepped with the fused gather
base = MultiHoverAviary(num_envs=sh.count, **kw)      # stepped plainly, gathered with NCCL
gather = ObsGather(env, sh, learner=0)
env.reset(); base.reset()
g = torch.Generator(device=dev).manual_seed(100 + rank)
acts = [torch.rand((sh.count, D, 4), device=dev, generator=g) * 2 - 1 for _ in range(8)]
out = {"world": world, "drones_per_rank": E_per * D}
# ---- correctness: the learner's tensors == NCCL all-gather of the plain step, bit for bit --------------------------------------
ok = True
for t in range(12):
    o, r, te, tr, _ = env.step(acts[t % 8])
    res = gather.wait()
    o2, r2, te2, tr2, _ = base.step(acts[t % 8])
    go = all_gather_envs(o2, sh); gr = all_gather_envs(r2, sh); gt = all_gather_envs(te2.to(torch.uint8), sh); gu = all_gather_envs(tr2.to(torch.uint8), sh)
    torch.cuda.synchronize()
    if rank == 0:
        ok &= bool(torch.equal(res[0], go)) and bool(torch.equal(res[1], gr)) and bool(torch.equal(res[2].to(torch.uint8), gt)) and bool(torch.equal(res[3].to(torch.uint8), gu))
    dist.barrier()
out["fused_equals_nccl_all_gather"] = ok if rank == 0 else None
out["timed_out"] = gather.timed_out() if rank == 0 else None


def timed(fn, n=50, reps=3):
    best = []
    for _ in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(n):
            fn(k)
        e1.record(); torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1) / n)
    t = torch.tensor([min(best)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def fused(k):
    env.step(acts[k % 8])
    gather.wait()


def nccl(k):
    o, *_ = base.step(acts[k % 8])
    all_gather_envs(o, sh)


def plain(k):
    base.step(acts[k % 8])


out["ms_per_tick_step_only"] = timed(plain)
out["ms_per_tick_fused_gather"] = timed(fused)
out["ms_per_tick_step_plus_nccl_all_gather_obs"] = timed(nccl)
bytes_in = (world - 1) * E_per * D * 72 * 4
out["learner_inbound_bytes_per_tick"] = bytes_in
out["fused_inbound_GBps"] = bytes_in / (out["ms_per_tick_fused_gather"] * 1e-3) / 1e9
out["nccl_inbound_GBps"] = bytes_in / (out["ms_per_tick_step_plus_nccl_all_gather_obs"] * 1e-3) / 1e9
if rank == 0:
    try:
        txt = open(os.environ["NCCL_DEBUG_FILE"]).read()
        out["nccl_transport_lines"] = [l.split("NCCL INFO ")[-1] for l in txt.splitlines() if " via " in l or "NVLS" in l or "P2P" in l][:12]
    except Exception as ex:
        out["nccl_transport_lines"] = repr(ex)
    print(json.dumps(out))
dist.destroy_process_group()
