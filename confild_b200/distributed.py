"""Frame-sharded decode across the GPUs of one box (one process per GPU, torch.distributed).

Every (frame, point) pair of the decoder is independent (reference: cnf/nf_networks.py:491-494 has no
cross-frame or cross-point operation), so frames are split across ranks with no data-path
collective; the only exchange is one all-gather of the decoded field.  The DPS
gradient dL/dlatent stays rank-local (each rank owns the rows of its own frames).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(T: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced split of ``T`` frames: the first ``T % world`` ranks get one extra."""
    base, rem = divmod(int(T), int(world))
    bounds, start = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        bounds.append((start, start + n))
        start += n
    return bounds


def local_frames(latents: torch.Tensor, group=None) -> Tuple[torch.Tensor, Tuple[int, int]]:
    """This rank's slice of ``latents (T, ...)`` and its ``(start, end)`` frame range."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    s, e = shard_bounds(latents.shape[0], world)[rank]
    return latents[s:e], (s, e)


def all_gather_frames(local_out: torch.Tensor, T: int, group=None) -> torch.Tensor:
    """Gather per-rank ``(T_r, ...)`` blocks (split as in ``shard_bounds``) into ``(T, ...)``.

    Equal shards use a single ``all_gather_into_tensor`` (NCCL over NVLink on GPUs); ragged shards
    are padded to the largest shard for the collective and trimmed afterwards.
    """
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_out
    world = dist.get_world_size(group)
    bounds = shard_bounds(T, world)
    sizes = [e - s for s, e in bounds]
    tail = tuple(local_out.shape[1:])
    tmax = max(sizes)
    if min(sizes) == tmax:
        full = local_out.new_empty((T,) + tail)
        dist.all_gather_into_tensor(full, local_out.contiguous(), group=group)
        return full
    padded = local_out.new_zeros((tmax,) + tail)
    padded[: local_out.shape[0]] = local_out
    gathered = local_out.new_empty((world * tmax,) + tail)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    return torch.cat([gathered[r * tmax: r * tmax + sizes[r]] for r in range(world)], dim=0)


def decode_frame_sharded(decode_fn: Callable[[torch.Tensor], torch.Tensor], latents: torch.Tensor,
                         group=None, gather: bool = True, chunks: int = 1) -> torch.Tensor:
    """Decode ``latents (T, L)`` with frames sharded over the ranks of ``group``.

    ``decode_fn(latents_local (T_r, L)) -> (T_r, P, cout)`` is the local decode (the CUDA module on
    GPUs).  With ``gather`` the decoded field is all-gathered to every rank.  ``chunks > 1`` splits the
    local frames so that the all-gather of chunk i (issued asynchronously) overlaps the decode of
    chunk i+1; the result is bit-identical to ``chunks == 1``.
    """
    T = latents.shape[0]
    lat_local, _ = local_frames(latents, group)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if not gather or world == 1:
        return decode_fn(lat_local)
    sizes = [e - s for s, e in shard_bounds(T, world)]
    if chunks <= 1 or min(sizes) != max(sizes) or sizes[0] % chunks != 0:
        return all_gather_frames(decode_fn(lat_local), T, group)
    # chunked, overlapped all-gather: every rank contributes frames [c*step, (c+1)*step) of its shard
    step = sizes[0] // chunks
    works, parts = [], []
    full: Optional[torch.Tensor] = None
    for c in range(chunks):
        part = decode_fn(lat_local[c * step:(c + 1) * step]).contiguous()
        if full is None:
            full = part.new_empty((chunks, world, step) + tuple(part.shape[1:]))
        parts.append(part)
        works.append(dist.all_gather_into_tensor(full[c].reshape((world * step,) + tuple(part.shape[1:])), part,
                                                 group=group, async_op=True))
    for w in works:
        w.wait()
    # (chunk, rank, frame) -> (rank, chunk, frame) = global frame order
    return full.permute(1, 0, 2, *range(3, full.dim())).reshape((T,) + tuple(full.shape[3:]))


class FusedGatherDecoder:
    """Frame-sharded decode whose all-gather is fused into the decode kernel (GPUs of one NVLink/NVSwitch box).

    Every rank allocates the gathered field ``(world*T_local, P, cout)`` in torch symmetric memory and maps its peers'
    buffers; ``__call__`` then runs ``cnf_forward_gather``: the kernel's epilogue stores each decoded point straight
    into all ``world`` buffers (12 bytes per point and target over NVLink), so there is no separate collective pass --
    only a device-side barrier after the stores.  Equal shards only (``T_local`` frames per rank).

    Cross-rank ordering.  The trailing barrier of call k tells every rank that all stores of call k have landed.  It
    does NOT stop a fast rank from starting call k+1 while a slow peer is still reading call k's result, so the
    symmetric allocation is double-buffered (``buffers=2``): call k writes buffer ``k % 2``.  A rank can only reach
    call k+2 (which overwrites buffer ``k % 2`` again) after passing the barrier of call k+1, i.e. after every peer has
    ENQUEUED call k+1 -- and a peer's reads of call k's result precede its call k+1 in stream order.  Contract: consume
    (or copy) the returned tensor on the calling stream before the next-but-one call.  ``buffers=1`` keeps a single
    buffer and adds a leading barrier instead (half the memory, one more barrier per call).
    """

    def __init__(self, model, T_local: int, P: int, group=None, buffers: int = 2):
        import torch.distributed._symmetric_memory as symm_mem

        if not dist.is_initialized():
            raise RuntimeError("FusedGatherDecoder needs an initialised process group")
        if buffers not in (1, 2):
            raise ValueError("buffers must be 1 or 2")
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("the fused gather supports up to 8 ranks (one NVSwitch box)")
        self.model, self.T_local, self.P = model, int(T_local), int(P)
        cout = int(model.net1[-1].weight.shape[0])
        dev = model.net1[0].weight.device
        self.buffers = buffers
        # one symmetric allocation holding `buffers` gathered fields back to back (one rendezvous, one signal pad)
        self.buf = symm_mem.empty((buffers, self.world * self.T_local, self.P, cout), dtype=torch.float32, device=dev)
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        block_bytes = self.T_local * self.P * cout * 4
        field_bytes = self.world * block_bytes
        # every rank's buffer b, offset to THIS rank's frame range
        self.out_ptrs = [[int(p) + b * field_bytes + self.rank * block_bytes for p in self.handle.buffer_ptrs]
                         for b in range(buffers)]
        self.calls = 0

    def __call__(self, coords: torch.Tensor, latents_local: torch.Tensor) -> torch.Tensor:
        b = self.calls % self.buffers
        self.calls += 1
        if self.buffers == 1:
            self.handle.barrier()  # every peer has finished reading the previous result (stream-ordered)
        self.model.decode_into(coords, latents_local, self.out_ptrs[b], T_expected=self.T_local)
        self.handle.barrier()  # all ranks' stores have landed (stream-ordered device barrier over the signal pads)
        return self.buf[b]
