"""One big formation sharded across GPUs.

Aviaries are independent, so the step path normally needs no communication (sharding.py).  The one exception is a
single formation larger than one GPU should hold: `_downwash` (BaseAviary.py:785-811) couples every drone to every
drone above it, so each physics substep needs the positions of the WHOLE formation: one exchange step per substep.

`FormationShard` is a `CtrlAviary` over this rank's contiguous slice of the formation.  Per substep it

  1. publishes its positions (`qs_dw_publish`): one kernel pushes the slice -- and the bounding boxes of its 32-drone
     chunks, which the force kernel uses to skip chunks that cannot contribute -- into the gathered array of every
     rank, its own and, through NVLink peer mappings, the peers', and then raises this rank's sequence flag on every
     rank (release store after a system fence);
  2. evaluates the downwash of its rows against the gathered array (`qs_downwash_rows`), whose kernel first waits
     (acquire loads, bounded spin) until every rank's flag carries the current sequence number;
  3. advances its drones by one substep (`qs_dyn_substeps` with the external force).

No host synchronisation, no NCCL call on the path (`exchange="p2p"`).  The gathered array is double buffered, which is
sufficient: a rank can publish sequence k+2 only after its own downwash k+1 has seen every peer's flag k+1, and a peer
raises k+1 only after its downwash k -- the last reader of buffer k & 1 -- has finished.  `exchange="nccl"` replaces
1. by `all_gather_into_tensor` (the baseline the p2p path is measured against, and the path the gloo CPU tests
cover); `exchange="local"` is the single-GPU case.

Lifetime and limits: the shards of one formation are created and destroyed collectively (a peer's mapping of this
rank's exchange allocation stays open for the life of the process); `exchange="p2p"` needs a drone count that is a
multiple of 32 x world; the p2p stage is not CUDA-graph capturable (the sequence number is a launch parameter).
"""
import ctypes as C

import os

import numpy as np
import torch
import torch.distributed as dist

from . import _native as N
from .envs.CtrlAviary import CtrlAviary
from .sharding import all_gather_envs, shard_envs
from .utils.enums import Physics


_IPC_IMPORTS = {}      # (importing device, 64-byte handle) -> mapped base pointer: a handle is opened once per process


def morton_order(xy, bits=16):
    """Permutation that sorts drones along a Z-order curve of their initial xy positions.  The downwash kernel skips
    32-drone chunks whose bounding boxes cannot interact, so index-coherent formations cost O(N k) instead of O(N^2);
    apply `initial_xyzs[morton_order(initial_xyzs[:, :2])]` when the formation's own order is not spatially coherent."""
    xy = np.asarray(xy, dtype=np.float64)
    lo, hi = xy.min(axis=0), xy.max(axis=0)
    q = ((xy - lo) / np.maximum(hi - lo, 1e-12) * ((1 << bits) - 1)).astype(np.uint64)

    def spread(v):
        v = v & np.uint64(0xFFFF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF)
        v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F)
        v = (v | (v << np.uint64(2))) & np.uint64(0x33333333)
        v = (v | (v << np.uint64(1))) & np.uint64(0x55555555)
        return v

    return np.argsort(spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)), kind="stable")


class FormationShard(CtrlAviary):
    """This rank's slice of one formation of `len(initial_xyzs)` drones (reference semantics: CtrlAviary with a
    downwash-enabled physics mode; actions are the RPMs of the LOCAL drones, observations their 20-float states)."""

    _EXTERNAL_DOWNWASH = True

    def __init__(self, initial_xyzs, initial_rpys=None, physics=Physics.PYB_DW, exchange="p2p", rank=None, world=None,
                 group=None, **kwargs):
        full = np.asarray(initial_xyzs, dtype=np.float64)
        if full.ndim != 2 or full.shape[1] != 3:
            raise ValueError("initial_xyzs must be [N_total, 3]")
        if exchange not in ("p2p", "nccl", "local"):
            raise ValueError("exchange must be 'p2p', 'nccl' or 'local'")
        self.shard = shard_envs(full.shape[0], rank, world)
        if exchange == "local" and self.shard.world != 1:
            raise ValueError("exchange='local' needs world == 1")
        if self.shard.world > N.MAX_PEERS:
            raise ValueError("at most %d ranks" % N.MAX_PEERS)
        if exchange == "p2p" and full.shape[0] % (32 * self.shard.world):
            raise ValueError("exchange='p2p' needs a drone count that is a multiple of 32 x world (a 32-drone chunk never "
                             "straddles ranks); pad the formation or use exchange='nccl'")
        self.exchange = exchange
        self.group = group
        self.N_TOTAL = int(full.shape[0])
        sl = slice(self.shard.start, self.shard.stop)
        rp = None if initial_rpys is None else np.asarray(initial_rpys, dtype=np.float64)[sl]
        kwargs.pop("num_drones", None)
        kwargs.setdefault("num_envs", 1)
        super().__init__(num_drones=self.shard.count, initial_xyzs=full[sl], initial_rpys=rp, physics=physics, **kwargs)
        if not (self._effects & N.EFFECT_DW):
            raise ValueError("FormationShard needs a physics mode with downwash (PYB_DW, PYB_GND_DRAG_DW)")
        dev = self.device
        # ONE exchange allocation per rank: [gathered array (positions + chunk boxes), buffer 0 | buffer 1 | flag words]
        nf = self._nf = int(self._lib.qs_dw_gathered_floats(self.N_TOTAL))
        self._xbuf = torch.zeros((2 * nf + N.MAX_PEERS,), dtype=torch.float32, device=dev)
        self._gathered = [self._xbuf[b * nf:(b + 1) * nf] for b in range(2)]
        self._flags_dev = self._xbuf[2 * nf:].view(torch.int32)
        self._counter = torch.zeros((1,), dtype=torch.int32, device=dev)
        self._err = torch.zeros((1,), dtype=torch.int32, device=dev)
        self._seq = 0
        self._pub_seq = -1            # sequence number under which the CURRENT positions are already in every rank's buffer (fused publish)
        self._fuse_publish = exchange == "p2p" and os.environ.get("QS_FUSED_PUBLISH", "1") != "0"
        self._peer_keepalive = None
        self._gathered_ptrs = None
        self._flag_ptrs = None
        if exchange == "p2p":
            if self.shard.world == 1:
                self.connect([self])
            elif dist.is_available() and dist.is_initialized():
                self._connect_ipc()

    # ---- peer wiring ------------------------------------------------------------------------------------------------
    def connect(self, shards):
        """Wires the peers of an in-process formation (one FormationShard per rank, any devices of this process)."""
        w = self.shard.world
        if len(shards) != w or any(s.shard.rank != r for r, s in enumerate(shards)):
            raise ValueError("connect() needs the %d shards in rank order" % w)
        with self._on_device():
            for s in shards:
                if s.device != self.device:
                    N.check(self._lib.qs_enable_peer_access(s.device.index), "qs_enable_peer_access")
        self._set_peer_bases([s._xbuf.data_ptr() for s in shards], keepalive=[s._xbuf for s in shards])

    def _connect_ipc(self):
        """One process per GPU: every rank's exchange allocation is mapped into every peer through CUDA IPC handles
        (qs_ipc_export / qs_ipc_import; peer access is enabled by the import) passed over the process group."""
        w, r = self.shard.world, self.shard.rank
        bases = [None] * w
        with self._on_device():
            handle, off = (C.c_ubyte * 64)(), C.c_ulonglong(0)
            N.check(self._lib.qs_ipc_export(self._xbuf.data_ptr(), handle, C.byref(off)), "qs_ipc_export")
            every = [None] * w
            dist.all_gather_object(every, (bytes(handle), int(off.value)), group=self.group)
            for q in range(w):
                if q == r:
                    bases[q] = self._xbuf.data_ptr()
                    continue
                hq, oq = every[q]
                key = (self.device.index, hq)
                if key not in _IPC_IMPORTS:
                    p = C.c_void_p()
                    N.check(self._lib.qs_ipc_import((C.c_ubyte * 64).from_buffer_copy(hq), 0, C.byref(p)), "qs_ipc_import")
                    _IPC_IMPORTS[key] = p.value
                bases[q] = _IPC_IMPORTS[key] + oq
        self._set_peer_bases(bases)
        dist.barrier(group=self.group)

    def _set_peer_bases(self, bases, keepalive=None):
        w, nf = self.shard.world, self._nf
        self._peer_keepalive = keepalive
        self._gathered_ptrs = [(C.c_void_p * w)(*[b + 4 * k * nf for b in bases]) for k in range(2)]
        self._flag_ptrs = (C.c_void_p * w)(*[b + 4 * 2 * nf for b in bases])
        pub = N.QsDwPublish()
        pub.flags, pub.counter = self._flag_ptrs, self._counter.data_ptr()
        pub.n_total, pub.world, pub.rank, pub.offset = self.N_TOTAL, w, self.shard.rank, self.shard.start
        self._pub = pub

    # ---- the exchange step ------------------------------------------------------------------------------------------
    def _downwash_stage(self, stream):
        L, sh = self._lib, self.shard
        if self.exchange == "local":
            return super()._downwash_stage(stream)
        rows = self._pos_f32          # float32 position mirror, refreshed by every substep kernel
        self._seq += 1
        buf = self._gathered[self._seq & 1]
        if self.exchange == "p2p":
            if self._gathered_ptrs is None:
                raise RuntimeError("FormationShard(exchange='p2p'): peers are not connected (connect() / process group)")
            if self._pub_seq != self._seq:      # not already pushed by the previous dynamics launch (first stage after a reset, ...)
                N.check(L.qs_dw_publish(rows.data_ptr(), sh.count, sh.start, self._gathered_ptrs[self._seq & 1], self.N_TOTAL, self._flag_ptrs,
                                        sh.world, sh.rank, self._seq, self._counter.data_ptr(), stream), "qs_dw_publish")
            N.check(L.qs_downwash_rows(C.byref(self._P), rows.data_ptr(), sh.count, buf.data_ptr(), self.N_TOTAL,
                                       self._flags_dev.data_ptr(), self._seq, sh.world, self._err.data_ptr(),
                                       self._dw_fz.data_ptr(), stream), "qs_downwash_rows")
            return
        self._all_gather_positions(rows, buf[:4 * self.N_TOTAL].view(self.N_TOTAL, 4))
        N.check(L.qs_dw_boxes(buf.data_ptr(), self.N_TOTAL, stream), "qs_dw_boxes")
        N.check(L.qs_downwash_rows(C.byref(self._P), rows.data_ptr(), sh.count, buf.data_ptr(), self.N_TOTAL, None, 0, 0, None,
                                   self._dw_fz.data_ptr(), stream), "qs_downwash_rows")

    def _dyn_substep(self, rpm_ptr, state20_ptr, flags, stream):
        """The dynamics launch also pushes the new positions (and chunk boxes, and this rank's flag) to every rank, under the
        sequence number of the NEXT downwash stage: the exchange costs no launch of its own (qs_dyn_substeps_pub)."""
        if not self._fuse_publish or self._gathered_ptrs is None:
            return super()._dyn_substep(rpm_ptr, state20_ptr, flags, stream)
        sh, nxt = self.shard, self._seq + 1
        pub = self._pub
        pub.gathered, pub.seq = self._gathered_ptrs[nxt & 1], nxt
        rc = self._lib.qs_dyn_substeps_pub(C.byref(self._P), C.byref(self._st), rpm_ptr, state20_ptr, self._dw_fz.data_ptr(),
                                           self._E, self._D, 1, self._effects, flags, C.byref(pub), stream)
        self._pub_seq = nxt
        return rc

    def reset(self, *args, **kwargs):
        self._pub_seq = -1            # the reset positions have not been pushed
        return super().reset(*args, **kwargs)

    def reorder_by_morton(self, *args, **kwargs):
        self._pub_seq = -1
        return super().reorder_by_morton(*args, **kwargs)

    def _all_gather_positions(self, rows, out):
        all_gather_envs(rows, self.shard, group=self.group, out=out)

    def exchange_timed_out(self):
        """True if a downwash kernel gave up waiting for a peer's positions (synchronises)."""
        return bool(self._err.item())
