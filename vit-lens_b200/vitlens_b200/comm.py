"""Symmetric peer memory over NVLink / NVSwitch (csrc/comm.cu): one arena per rank, mapped into every rank of the box.

What runs over it (one process per GPU, all ranks on one node):
  * the contrastive loss reads the other ranks' feature blocks IN PLACE inside its logits GEMMs (ops.PeerRows -> vl_gemm_bf16
    b_peers): the all-gather of reference loss.py:55-76 is fused into the kernel; the small exchanges (row-LSE vectors, loss and
    d(scale) partial sums) are rank-ordered peer loads;
  * grad_sync.GradReducer pushes gradient buckets to every peer with the copy engines while backward runs, and the fused AdamW
    adds the world copies in rank order (no SM-resident collective next to the persistent GEMMs).
torch.distributed is used once, to exchange the 64-byte IPC handles.  VL_COMM=nccl disables the arena (NCCL collectives
instead).  Arena layout: [flags 2 KiB][small-exchange ring][feature ring][bump-allocated regions (gradients)]."""
from __future__ import annotations

import os
from typing import List, Optional

import torch

from . import lib as L

FLAG_BYTES = 64 * 8 * 4
FLAG_FEAT = 0            # tickets of the feature ring
FLAG_SMALL0 = 1          # .. +SMALL_SLOTS-1: small-exchange ring
SMALL_SLOTS = 4
SMALL_BYTES = 64 << 10
FEAT_SLOTS = 4
FLAG_GRAD0 = 8           # .. 63: one flag per gradient bucket
MAX_GRAD_FLAGS = 56

_ARENA: Optional["PeerArena"] = None


class _Alias:
    """__cuda_array_interface__ view of raw device memory (lets torch wrap arena memory without owning it)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


_TYPESTR = {torch.float32: "<f4", torch.bfloat16: "<u2", torch.int32: "<i4", torch.uint8: "|u1"}


class PeerArena:
    def __init__(self, rank: int, world: int, nbytes: int, feat_bytes: int = 16 << 20, group=None):
        import torch.distributed as dist

        self.rank, self.world, self.group = rank, world, group
        self.feat_bytes = (feat_bytes + 255) // 256 * 256
        self.small_off = FLAG_BYTES
        self.feat_off = self.small_off + SMALL_SLOTS * SMALL_BYTES
        self.cursor = self.feat_off + FEAT_SLOTS * self.feat_bytes
        self.nbytes = self.cursor + int(nbytes) + 4096  # `nbytes`: room for alloc() (gradient regions) on top of the fixed rings
        handle, self.base = L.comm_init(rank, world, self.nbytes)
        handles: List[bytes] = [b""] * world
        if world > 1:
            dist.all_gather_object(handles, handle, group=group)
        else:
            handles = [handle]
        L.comm_connect(b"".join(handles))
        self.peers = [L.comm_peer_ptr(p) for p in range(world)]
        self.tickets = {"feat": 0, "small": 0}
        self.device = torch.device("cuda", torch.cuda.current_device())
        if world > 1:
            dist.barrier(group=group)  # every rank has mapped every arena before anyone signals

    # ---- memory
    def alloc(self, nbytes: int) -> int:
        off = (self.cursor + 255) // 256 * 256
        if off + nbytes > self.nbytes:
            raise MemoryError(f"peer arena exhausted: need {nbytes} bytes at offset {off} of {self.nbytes}")
        self.cursor = off + nbytes
        return off

    def tensor(self, offset: int, shape, dtype) -> torch.Tensor:
        """A tensor aliasing THIS rank's arena at `offset` (bf16 goes through a uint16 view)."""
        t = torch.as_tensor(_Alias(self.base + offset, shape, _TYPESTR[dtype]), device=self.device)
        return t.view(torch.bfloat16) if dtype == torch.bfloat16 else t

    def peer_addr(self, peer: int, offset: int) -> int:
        return self.peers[peer] + offset

    def flag_addr(self, idx: int) -> int:
        return self.base + idx * 8 * 4

    # ---- feature ring (the fused all-gather of the contrastive loss)
    def publish_features(self, feats) -> list:
        """feats: k fp32 [B_loc, E] tensors.  Writes their bf16 copies into this rank's next ring slot, publishes the ticket and
        returns one ops.PeerRows per tensor (the [world * B_loc, E] matrix, rows of rank p living in p's arena)."""
        from . import ops

        k = len(feats)
        Bl, E = feats[0].shape
        need = k * Bl * E * 2
        if need > self.feat_bytes:
            raise MemoryError(f"feature block of {need} bytes exceeds the arena's feature slots ({self.feat_bytes} bytes)")
        self.tickets["feat"] += 1
        t = self.tickets["feat"]
        off = self.feat_off + (t % FEAT_SLOTS) * self.feat_bytes
        buf = self.tensor(off, (k, Bl, E), torch.bfloat16)
        for i, f in enumerate(feats):
            ops.cast_bf16(f.detach().float(), out=buf[i])
        L.allgather_features(off, need, FLAG_FEAT, t)
        return [ops.PeerRows(addrs=[self.peer_addr(p, off + i * Bl * E * 2) for p in range(self.world)], rows=Bl, E=E,
                             flags=self.flag_addr(FLAG_FEAT), ticket=t, local=buf[i], arena=self) for i in range(k)]

    def features_alive(self, ticket: int) -> bool:
        return self.tickets["feat"] - ticket < FEAT_SLOTS - 1

    # ---- small exchanges (fp32 vectors up to SMALL_BYTES): rank-ordered, deterministic
    def _small(self, vec: torch.Tensor):
        v = vec.detach().float().reshape(-1).contiguous()
        n = v.numel()
        assert n * 4 <= SMALL_BYTES, "small exchange too large"
        self.tickets["small"] += 1
        t = self.tickets["small"]
        slot = t % SMALL_SLOTS
        off = self.small_off + slot * SMALL_BYTES
        self.tensor(off, (n,), torch.float32).copy_(v)
        L.comm_signal(FLAG_SMALL0 + slot, t)
        L.comm_wait(FLAG_SMALL0 + slot, t)
        return off, n

    def all_gather_vec(self, vec: torch.Tensor) -> torch.Tensor:
        """[n] -> [world * n], rank-major (what dist.all_gather_into_tensor returns)."""
        off, n = self._small(vec)
        out = torch.empty((self.world * n,), device=vec.device, dtype=torch.float32)
        L.comm_peer_gather(off, n, out)
        return out

    def all_reduce_sum(self, vec: torch.Tensor) -> torch.Tensor:
        """Sum over ranks in rank order (bit-identical on every rank)."""
        off, n = self._small(vec)
        out = torch.empty((n,), device=vec.device, dtype=torch.float32)
        L.comm_peer_reduce(off, n, out)
        return out.reshape(vec.shape)


def arena() -> Optional[PeerArena]:
    return _ARENA


def init_arena(nbytes: int = 0, feat_bytes: int = 16 << 20, group=None) -> Optional[PeerArena]:
    """Create the process-wide arena (idempotent): the flag block, the small-exchange and feature rings plus `nbytes` of
    bump-allocated space (GradReducer needs world x 4 bytes x trainable parameters).  Returns None when peer memory is switched
    off (VL_COMM=nccl) or torch.distributed is not initialised with more than one rank."""
    global _ARENA
    import torch.distributed as dist

    if _ARENA is not None:
        return _ARENA
    if os.environ.get("VL_COMM", "").lower() == "nccl":
        return None
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2 or not torch.cuda.is_available():
        return None
    _ARENA = PeerArena(dist.get_rank(group), dist.get_world_size(group), nbytes, feat_bytes, group)
    return _ARENA


def destroy_arena():
    global _ARENA
    if _ARENA is not None:
        torch.cuda.synchronize()
        L.comm_destroy()
        _ARENA = None
