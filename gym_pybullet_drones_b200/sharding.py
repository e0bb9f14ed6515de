"""Multi-GPU layout of the aviaries: one process per GPU, each owning a contiguous slice of the aviary axis.

The step path needs NO communication -- every aviary is independent (the only couplings in the reference,
`_downwash` and MultiHover's reward/termination reductions, stay inside one aviary, BaseAviary.py:785-811,
MultiHoverAviary.py:75-130), and an aviary never straddles GPUs.  The only collective offered is an optional
all-gather of observations / rewards / flags for a learner that wants single tensors (torch.distributed:
NCCL over NVLink on GPUs, gloo on CPU for the tests).
"""
from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class EnvShard:
    """Aviaries [start, stop) of `total` belong to `rank` of `world`."""
    rank: int
    world: int
    total: int
    start: int
    stop: int

    @property
    def count(self):
        return self.stop - self.start


def shard_envs(total_envs: int, rank: int = None, world: int = None) -> EnvShard:
    """Contiguous, balanced partition of the aviary axis (the first `total % world` ranks get one extra)."""
    if world is None:
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    if total_envs < world:
        raise ValueError("fewer aviaries (%d) than ranks (%d)" % (total_envs, world))
    base, extra = divmod(total_envs, world)
    start = rank * base + min(rank, extra)
    return EnvShard(rank, world, total_envs, start, start + base + (1 if rank < extra else 0))


def all_gather_envs(local: torch.Tensor, shard: EnvShard, group=None, out: torch.Tensor = None) -> torch.Tensor:
    """All-gathers a per-aviary tensor [E_local, ...] into [E_total, ...] (same order as a single-GPU run).
    Off the step path: call it only where the learner needs one tensor.  (formation.py uses it with `out=` for the
    positions of a formation sharded by drones: same partition, the unit is then a drone.)"""
    if shard.world == 1:
        if out is None:
            return local
        out.copy_(local)
        return out
    counts = [shard_envs(shard.total, r, shard.world).count for r in range(shard.world)]
    if len(set(counts)) == 1:
        if out is None:
            out = torch.empty((shard.total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # uneven shards: collectives need equal sizes, so pad every shard to the largest and trim after the gather
    out_arg = out
    m = max(counts)
    padded = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty((shard.world * m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    res = torch.cat([out[r * m:r * m + c] for r, c in enumerate(counts)], dim=0)
    if out_arg is not None:
        out_arg.copy_(res)
        return out_arg
    return res


class ObsGather:
    """Fused observation gather: instead of an all-gather AFTER the step, every rank's step kernel writes its
    finished observation rows -- and its per-aviary reward / terminated / truncated -- straight into the LEARNER rank's
    `[E_total, D, obs_dim]` tensor through a peer mapping of that tensor (CUDA IPC, NVLink), then raises its flag word on the
    learner (release at system scope, last warp of the grid).  The learner queues `wait()` (a one-warp kernel that acquires
    the `world` flags) on its stream before consuming `obs`; no NCCL call, no separate copy kernel, no host involvement.

        gather = ObsGather(env, shard)           # collective: every rank, after dist.init_process_group
        obs, rew, term, trunc, _ = env.step(a)   # also fills gather.obs / .reward / .terminated / .truncated on the learner
        gather.wait()                            # learner only (no-op elsewhere): orders the learner's stream after all ranks

    Single-process use (tests, one GPU or several visible devices): `ObsGather.connect_local([gather_0, gather_1, ...])`.
    The destination is double buffered by the tick's parity, so a rank may run one tick ahead of the learner."""

    def __init__(self, env, shard: EnvShard, learner: int = 0, group=None, local=False):
        import ctypes as C
        from . import _native as N
        self.env, self.shard, self.learner, self.group = env, shard, learner, group
        self._lib, self._N, self._C = N.lib(), N, C
        E_tot, D, od, dev = shard.total, env._D, env._obs_dim, env.device
        self.is_learner = shard.rank == learner
        self.seq = 0
        self._counter = torch.zeros((1,), dtype=torch.int32, device=dev)
        self._err = torch.zeros((1,), dtype=torch.int32, device=dev)
        if self.is_learner:
            # ONE allocation (one IPC handle): 2 x [obs | reward | terminated | truncated] + flag words
            self._sizes = self._layout(E_tot, D, od)
            self._buf = torch.zeros((self._sizes["total"],), dtype=torch.uint8, device=dev)
            self._views(self._buf)
        self._base = None
        if not local:
            self._connect_ipc()

    @staticmethod
    def _layout(E, D, od):
        a = lambda n: (n + 255) // 256 * 256      # noqa: E731
        obs, rew, flg = a(E * D * od * 4), a(E * 4), a(E)
        per = obs + rew + 2 * flg
        return dict(obs=obs, rew=rew, flg=flg, per=per, flags_off=2 * per, total=2 * per + 256)

    def _views(self, buf):
        s, sh, env = self._sizes, self.shard, self.env
        self.obs, self.reward, self.terminated, self.truncated = [], [], [], []
        for b in range(2):
            o = b * s["per"]
            self.obs.append(buf[o:o + sh.total * env._D * env._obs_dim * 4].view(torch.float32).view(sh.total, env._D, env._obs_dim))
            self.reward.append(buf[o + s["obs"]:o + s["obs"] + sh.total * 4].view(torch.float32))
            self.terminated.append(buf[o + s["obs"] + s["rew"]:o + s["obs"] + s["rew"] + sh.total].view(torch.bool))
            self.truncated.append(buf[o + s["obs"] + s["rew"] + s["flg"]:o + s["obs"] + s["rew"] + s["flg"] + sh.total].view(torch.bool))
        self.flags = buf[s["flags_off"]:s["flags_off"] + 64].view(torch.int32)

    def _connect_ipc(self):
        C, N = self._C, self._N
        sh = self.shard
        if sh.world == 1:
            return self._set_base(self._buf.data_ptr())
        payload = [None]
        if self.is_learner:
            with torch.cuda.device(self.env.device):
                handle, off = (C.c_ubyte * 64)(), C.c_ulonglong(0)
                N.check(self._lib.qs_ipc_export(self._buf.data_ptr(), handle, C.byref(off)), "qs_ipc_export")
            payload = [(bytes(handle), int(off.value))]
        dist.broadcast_object_list(payload, src=self.learner, group=self.group)
        if self.is_learner:
            base = self._buf.data_ptr()
        else:
            hq, oq = payload[0]
            p = C.c_void_p()
            with torch.cuda.device(self.env.device):
                N.check(self._lib.qs_ipc_import((C.c_ubyte * 64).from_buffer_copy(hq), 0, C.byref(p)), "qs_ipc_import")
            base = p.value + oq
        self._set_base(base)
        dist.barrier(group=self.group)

    @staticmethod
    def connect_local(gathers):
        """In-process wiring: `gathers` = one ObsGather(local=True) per rank; peer access is enabled between their devices."""
        lead = next(g for g in gathers if g.is_learner)
        for g in gathers:
            if g.env.device != lead.env.device:
                with torch.cuda.device(g.env.device):
                    g._N.check(g._lib.qs_enable_peer_access(lead.env.device.index), "qs_enable_peer_access")
            g._keep = lead._buf
            g._sizes = lead._sizes
            g._set_base(lead._buf.data_ptr())

    def _set_base(self, base):
        sh, env = self.shard, self.env
        self._sizes = getattr(self, "_sizes", None) or self._layout(sh.total, env._D, env._obs_dim)
        self._base = base
        env._gather = self

    def _io_fields(self, parity):
        """(obs, reward, terminated, truncated, flag) destination pointers of THIS rank's rows for a tick of that parity."""
        s, sh, env = self._sizes, self.shard, self.env
        o = self._base + parity * s["per"]
        return (o + sh.start * env._D * env._obs_dim * 4, o + s["obs"] + sh.start * 4, o + s["obs"] + s["rew"] + sh.start,
                o + s["obs"] + s["rew"] + s["flg"] + sh.start, self._base + s["flags_off"] + 4 * sh.rank)

    def arm(self, io):
        """Called by the env right before a tick is launched: points the QsStepIO at this tick's destination."""
        self.seq += 1
        io.obs_gather, io.reward_gather, io.terminated_gather, io.truncated_gather, io.gather_flag = self._io_fields(self.seq & 1)
        io.gather_counter, io.gather_seq = self._counter.data_ptr(), self.seq

    def wait(self):
        """Learner: orders the current stream after every rank's flag of the last armed tick; returns the gathered views
        (obs [E_total, D, obs_dim], reward, terminated, truncated).  Other ranks: returns None."""
        if not self.is_learner:
            return None
        with torch.cuda.device(self.env.device):
            self._N.check(self._lib.qs_wait_flags(self.flags.data_ptr(), self.seq, self.shard.world, self._err.data_ptr(),
                                                  torch.cuda.current_stream(self.env.device).cuda_stream), "qs_wait_flags")
        b = self.seq & 1
        return self.obs[b], self.reward[b], self.terminated[b], self.truncated[b]

    def timed_out(self):
        return bool(self._err.item())
