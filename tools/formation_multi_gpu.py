"""One formation sharded across the GPUs of a node.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 \
        tools/formation_multi_gpu.py [--drones 16384] [--ticks 10]

Checks that the sharded run (exchange = NCCL all-gather, and the push + flag kernels over NVLink peer memory) is
bit-identical to the unsharded formation on rank 0, then times one exchange + downwash stage and one whole control tick per mode
(CUDA events, max over ranks).  Writes gpurun_out/formation_multi_gpu.json on rank 0.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_pybullet_drones_b200.formation import FormationShard, morton_order  # noqa: E402
from gym_pybullet_drones_b200.sharding import all_gather_envs  # noqa: E402
from gym_pybullet_drones_b200.utils.enums import Physics  # noqa: E402


def stacks(nx, ny, pitch=1.6):
    k = np.arange(nx * ny * 4)
    st, ly = k // 4, k % 4
    return np.stack([pitch * (st % nx) + 0.04 * ly, pitch * (st // nx) - 0.03 * ly, 0.5 + 1.5 * ly], axis=1).astype(np.float32).astype(np.float64)


def config4(n):
    side = int(round(n ** 0.5))
    i = np.arange(side * side)
    xyz = np.stack([0.15 * (i % side), 0.15 * (i // side), 0.1 + 0.05 * (i % 16)], axis=1)
    return xyz[morton_order(xyz[:, :2])]


LOG = None


def log(*a):
    print(*a, file=LOG, flush=True)


def main():
    global LOG
    import faulthandler
    import traceback
    ap = argparse.ArgumentParser()
    ap.add_argument("--drones", type=int, default=16384)
    ap.add_argument("--ticks", type=int, default=10)
    ap.add_argument("--timing-only", action="store_true")
    ap.add_argument("--modes", default="nccl,p2p", help="exchange modes to check and time")
    ap.add_argument("--tag", default="", help="suffix of the output file (e.g. 'unfused' for a QS_FUSED_PUBLISH=0 run)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    os.makedirs("gpurun_out", exist_ok=True)
    LOG = open("gpurun_out/formation_w%d_n%d_rank%d.log" % (world, args.drones, rank), "w")
    faulthandler.enable(LOG)
    try:
        run(args, rank, world, local)
    except BaseException:
        log(traceback.format_exc())
        raise


def run(args, rank, world, local):
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    saved = os.dup(1)
    os.dup2(2, 1)                                    # NCCL prints its banner on stdout
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier()
    os.dup2(saved, 1)
    log("alloc conf", os.environ.get("PYTORCH_CUDA_ALLOC_CONF"), "peer access", [torch.cuda.can_device_access_peer(local, d) for d in range(torch.cuda.device_count()) if d != local])
    side = int(round((args.drones // 4) ** 0.5))
    xyz = stacks(side, side)
    n = len(xyz)
    kw = dict(physics=Physics.PYB_GND_DRAG_DW, pyb_freq=240, ctrl_freq=48)
    rng = np.random.default_rng(21)
    hover = 14468.429183500699
    acts = [torch.from_numpy((hover * (1 + 0.05 * rng.uniform(-1, 1, (1, n, 4)))).astype(np.float32)).to(dev) for _ in range(args.ticks)]
    res = {"drones": n, "world": world, "ticks": args.ticks, "substeps_per_tick": 5}

    ref = None
    if rank == 0 and not args.timing_only:
        env = FormationShard(xyz, exchange="local", rank=0, world=1, **kw)
        env.reset()
        for a in acts:
            ref = env.step(a)[0]
        ref = ref.clone()
        del env
    out_path = "gpurun_out/formation_multi_gpu_%d_%d%s.json" % (world, args.drones, ("_" + args.tag) if args.tag else "")
    res["fused_publish"] = os.environ.get("QS_FUSED_PUBLISH", "1") != "0"
    log("reference done")
    modes = [m for m in args.modes.split(",") if m in ("nccl", "p2p")]
    for name, mode in (() if args.timing_only else tuple((m, m) for m in modes)):
        log("mode", name, "construct")
        env = FormationShard(xyz, exchange=mode, **kw)
        log("mode", name, "connected")
        sh = env.shard
        env.reset()
        dist.barrier()
        ok = True
        for t, a in enumerate(acts):
            obs = env.step(a[:, sh.start:sh.stop].contiguous())[0]
            if t == 0 and mode == "p2p" and env.exchange_timed_out():
                ok = False
                break
        torch.cuda.synchronize()
        timed_out = env.exchange_timed_out() if mode == "p2p" else False
        full = all_gather_envs(obs[0].contiguous(), sh)
        if rank == 0:
            res[name + "_bit_identical_to_unsharded"] = bool(ok and torch.equal(full, ref[0]))
            res[name + "_max_abs_diff"] = float((full - ref[0]).abs().max())
        res[name + "_timed_out"] = bool(timed_out)
        log("mode", name, "done", {k: v for k, v in res.items() if k.startswith(name)})
        if rank == 0:
            json.dump(res, open(out_path, "w"), indent=1)
        del env
        dist.barrier()

    # timing on the config-4 geometry (Morton order), reset positions: the whole stage per mode, and its parts
    import ctypes as C
    from gym_pybullet_drones_b200 import _native as N
    xyz4 = config4(args.drones)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    L = N.lib()

    def timed(fn, K=200):
        dist.barrier()
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / K], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) * 1e3

    for mode in ["local"] * (world == 1) + modes:
        log("timing", mode)
        env = FormationShard(xyz4, exchange=mode, **kw)
        env.reset()
        sh = env.shard
        res["stage_us_" + mode] = timed(lambda: env._downwash_stage(sp))
        # The whole control tick of this rank's slice (S substeps: dynamics + exchange + downwash each) on the STACKS geometry
        # (1.6 m pitch, 1.5 m between layers), hover RPM, 5 ticks from reset.  Not on the config-4 lattice: with 0.15 m between
        # drones of (initially exactly) equal height, the reference's downwash term alpha ~ 1/dz^2 (BaseAviary.py:803) is singular
        # as soon as the heights differ by rounding, the drones are thrown apart and the boxes stop culling -- a property of the
        # model, measured as 92 us -> 1.1 ms per stage after two substeps.
        env_t = FormationShard(xyz, exchange=mode, **kw)
        a_loc = torch.full((1, env_t.shard.count, 4), float(hover), dtype=torch.float32, device=dev)

        def ticks(k):
            for _ in range(k):
                env_t.step(a_loc)
        env_t.reset(); ticks(2); env_t.reset()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ticks(5); e1.record()
        torch.cuda.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1) / 5], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        res["tick_us_" + mode] = float(tt.item()) * 1e3
        env_t.reset()
        res["stage_stacks_us_" + mode] = timed(lambda: env_t._downwash_stage(sp))
        del env_t
        dist.barrier()
        if mode == "p2p":
            res["stage_p2p_timed_out"] = bool(env.exchange_timed_out())

            def publish_only():
                env._seq += 1
                N.check(L.qs_dw_publish(env._pos_f32.data_ptr(), sh.count, sh.start, env._gathered_ptrs[env._seq & 1], env.N_TOTAL,
                                        env._flag_ptrs, sh.world, sh.rank, env._seq, env._counter.data_ptr(), sp), "qs_dw_publish")
            res["part_us_publish_kernel"] = timed(publish_only)
        if mode == "nccl":
            rows, buf = env._pos_f32, env._gathered[0]
            pos_view = buf[:4 * env.N_TOTAL].view(env.N_TOTAL, 4)
            res["part_us_nccl_all_gather"] = timed(lambda: env._all_gather_positions(rows, pos_view))
            res["part_us_boxes_kernel"] = timed(lambda: N.check(L.qs_dw_boxes(buf.data_ptr(), env.N_TOTAL, sp), "boxes"))
            res["part_us_downwash_rows_kernel"] = timed(lambda: N.check(L.qs_downwash_rows(
                C.byref(env._P), rows.data_ptr(), sh.count, buf.data_ptr(), env.N_TOTAL, None, 0, 0, None, env._dw_fz.data_ptr(), sp), "rows"))
        del env
        dist.barrier()
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(res, open(out_path, "w"), indent=1)
        print(json.dumps(res))
    log("finished")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
