"""N>1 host logic on CPU: world_size-2 gloo processes partition the aviary axis and all-gather per-aviary tensors.
(The step path itself has no collective; the GPU-side shard equivalence is tests/test_gpu_parity.py::test_full_size...)"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gym_pybullet_drones_b200.sharding import all_gather_envs, shard_envs


def test_shard_partition_properties():
    for total in (2, 7, 64, 65536, 262144 + 3):
        for world in (1, 2, 4, 8):
            if total < world:
                continue
            shards = [shard_envs(total, r, world) for r in range(world)]
            assert shards[0].start == 0 and shards[-1].stop == total
            assert all(a.stop == b.start for a, b in zip(shards, shards[1:]))
            assert max(s.count for s in shards) - min(s.count for s in shards) <= 1
    with pytest.raises(ValueError):
        shard_envs(3, 0, 4)
    with pytest.raises(ValueError):
        shard_envs(8, 9, 8)


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = shard_envs(total)
        assert (sh.rank, sh.world) == (rank, world)
        # every aviary e carries a recognisable payload; the sharded oracle step must equal the slice of the global one
        from oracle.dyn_oracle import OracleAviary
        rng = np.random.default_rng(0)
        acts = rng.uniform(-1, 1, (5, total, 2, 4)).astype(np.float32)
        env = OracleAviary("multihover", sh.count, 2, act="rpm")
        env.reset()
        for t in range(5):
            obs, rew, term, trunc = env.step(acts[t, sh.start:sh.stop])
        g_obs = all_gather_envs(torch.from_numpy(obs), sh)
        g_rew = all_gather_envs(torch.from_numpy(rew), sh)
        g_flag = all_gather_envs(torch.from_numpy(trunc.astype(np.uint8)), sh)
        if rank == 0:
            ref = OracleAviary("multihover", total, 2, act="rpm")
            ref.reset()
            for t in range(5):
                o, r, te, tr = ref.step(acts[t])
            q.put((bool(np.array_equal(g_obs.numpy(), o)), bool(np.array_equal(g_rew.numpy(), r)),
                   bool(np.array_equal(g_flag.numpy().astype(bool), tr))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_world2_gloo_sharded_equals_single(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == (True, True, True)


def _formation_worker(rank, world, port, q):
    """A formation sharded by DRONES: downwash needs every position each substep (formation.py); the host-side
    exchange (partition offsets, uneven all-gather into a preallocated buffer) against the unsharded oracle."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import dyn_oracle as O
        P = O.OracleParams()
        k = np.arange(27)
        xyz = np.stack([1.6 * (k // 3) + 0.04 * (k % 3), -0.03 * (k % 3), 0.5 + 1.5 * (k % 3)], axis=1)
        n = len(xyz)
        sh = shard_envs(n)
        rpm = np.full((n, 4), P.HOVER_RPM * 1.02)

        def run(pos, rows, gather):
            quat = np.tile([0., 0., 0., 1.], (len(rows), 1)); vel = np.zeros((len(rows), 3)); rr = np.zeros((len(rows), 3))
            pos = pos[rows].copy()
            buf = torch.zeros((n, 3), dtype=torch.float64)
            for _ in range(12):
                every = gather(pos, buf)
                fz = O.downwash_body_z(P, every[None])[0][rows]
                pos, quat, vel, rr, _ = O.dynamics_substep(P, rpm[rows], pos, quat, vel, rr, effects=O.EFFECT_DW, dw_fz=fz)
            return pos

        mine = np.arange(sh.start, sh.stop)
        local = run(xyz, mine, lambda p, buf: all_gather_envs(torch.from_numpy(p), sh, out=buf).numpy())
        g = all_gather_envs(torch.from_numpy(local), sh)
        if rank == 0:
            ref = run(xyz, np.arange(n), lambda p, buf: p)
            q.put((bool(np.array_equal(g.numpy(), ref)), bool(np.abs(ref - xyz).max() > 1e-6)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_formation_sharded_by_drones():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_formation_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == (True, True)
