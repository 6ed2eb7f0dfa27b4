"""Frame-range sharding across ranks (one process per GPU).

Encode shards by whole segments of frames_per_segment packets, decode by packets;
both are byte-identical to the unsharded result by construction.  The only cross-rank step is
putting the per-rank packet blocks back in order, and only when one consumer wants one buffer:
  concat_packets_to(dst, ...)  the packet-offset scan over the ranks' byte totals plus one point-to-point copy
                               per rank straight into its slice of the destination buffer (NCCL send/recv =
                               NVLink P2P copies on a B200 box); no collective on the data path;
  gather_packets(...)          every rank gets everything (a size all-gather plus a padded byte all-gather).
Works with any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple


def plan_packet_shards(num_packets: int, world: int, frames_per_segment: int) -> List[Tuple[int, int]]:
    """Contiguous [first_packet, count) per rank, boundaries aligned to whole segments.

    frames_per_segment == 0 means the stream is one serial chain: it cannot be split, rank 0 gets it.
    """
    if world < 1:
        raise ValueError("world must be >= 1")
    if frames_per_segment == 0:
        return [(0, num_packets)] + [(num_packets, 0)] * (world - 1)
    segs = (num_packets + frames_per_segment - 1) // frames_per_segment
    out = []
    for r in range(world):
        s0 = segs * r // world
        s1 = segs * (r + 1) // world
        p0 = min(s0 * frames_per_segment, num_packets)
        p1 = min(s1 * frames_per_segment, num_packets)
        out.append((p0, p1 - p0))
    return out


def plan_frame_shards(num_sample_frames: int, frame_size: int, world: int, frames_per_segment: int) -> List[Tuple[int, int]]:
    """Per-rank [first_sample_frame, num_sample_frames) following plan_packet_shards."""
    num_packets = (num_sample_frames + frame_size - 1) // frame_size
    out = []
    for p0, n in plan_packet_shards(num_packets, world, frames_per_segment):
        a = min(p0 * frame_size, num_sample_frames)
        b = min((p0 + n) * frame_size, num_sample_frames)
        out.append((a, b - a))
    return out


def gather_packets(packets, sizes, group=None):
    """All ranks call this with their packet bytes (uint8 tensor) and sizes (int32 tensor); every rank
    gets back (all packets in rank order, all sizes).  Tensors stay on their device."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dev = packets.device
    meta = torch.tensor([packets.numel(), sizes.numel()], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    nbytes = [int(m[0]) for m in metas]
    ncount = [int(m[1]) for m in metas]
    pad_b, pad_c = max(max(nbytes), 1), max(max(ncount), 1)
    pb = torch.zeros(pad_b, dtype=torch.uint8, device=dev)
    pb[:packets.numel()] = packets
    pc = torch.zeros(pad_c, dtype=torch.int32, device=dev)
    pc[:sizes.numel()] = sizes.to(torch.int32)
    all_b = [torch.empty_like(pb) for _ in range(world)]
    all_c = [torch.empty_like(pc) for _ in range(world)]
    dist.all_gather(all_b, pb, group=group)
    dist.all_gather(all_c, pc, group=group)
    out_p = torch.cat([b[:n] for b, n in zip(all_b, nbytes)])
    out_s = torch.cat([c[:n] for c, n in zip(all_c, ncount)])
    return out_p, out_s


def concat_packets_to(dst: int, packets, sizes, group=None):
    """Rank-ordered concatenation of every rank's packet block on rank `dst`, with point-to-point copies only.

    Every rank calls this with its packet bytes (uint8 tensor) and sizes (int32 tensor).  Rank `dst` learns the
    other ranks' totals (two int64 each), runs the exclusive scan that places each block, and receives each
    block directly into its slice of the output -- under NCCL these are peer-to-peer copies over NVLink, one per
    rank, with no staging and no padding.  Returns (packets, sizes) on `dst`, None elsewhere."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = packets.device
    sizes = sizes.to(torch.int32)
    meta = torch.tensor([packets.numel(), sizes.numel()], dtype=torch.int64, device=dev)
    if rank != dst:
        dist.send(meta, dst, group=group)
        if packets.numel():
            dist.send(packets.contiguous(), dst, group=group)
        if sizes.numel():
            dist.send(sizes.contiguous(), dst, group=group)
        return None
    metas = []
    for r in range(world):
        m = meta.clone()
        if r != dst:
            dist.recv(m, r, group=group)
        metas.append((int(m[0]), int(m[1])))
    # exclusive scan of the byte / packet totals = where each rank's block starts
    out_p = torch.empty(sum(m[0] for m in metas), dtype=torch.uint8, device=dev)
    out_s = torch.empty(sum(m[1] for m in metas), dtype=torch.int32, device=dev)
    ob = oc = 0
    for r, (nb, nc) in enumerate(metas):
        if r == dst:
            out_p[ob:ob + nb] = packets
            out_s[oc:oc + nc] = sizes
        else:
            if nb:
                dist.recv(out_p[ob:ob + nb], r, group=group)
            if nc:
                dist.recv(out_s[oc:oc + nc], r, group=group)
        ob += nb
        oc += nc
    return out_p, out_s


# ---------------------------------------------------------------------------------------------------------------------
# One packet buffer shared by all ranks of a job (one process per GPU): the cross-GPU step of SURVEY 8e inside the
# encode call itself.  The home rank owns the buffer, the packet-size table and the 1 KB exchange block; the other ranks
# map them with CUDA IPC (alac_b200_ipc_*), and every rank's enc_assemble_kernel then stores its packets at their final
# offset there over NVLink (alac_b200_encode_placed).  torch.distributed is used once, at set-up, to hand the three IPC
# handles round; nothing on the data path is a collective.
# ---------------------------------------------------------------------------------------------------------------------
class _RawDevice:
    """A raw device pointer exposed through __cuda_array_interface__ (no ownership)."""

    def __init__(self, ptr: int, count: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 3}


def tensor_from_pointer(ptr: int, count: int, device, typestr: str = "|u1"):
    """torch view of `count` elements at device pointer `ptr` (local or peer-mapped memory)."""
    import torch
    return torch.as_tensor(_RawDevice(ptr, count, typestr), device=device)


def staging_slots(slot_bytes, align: int = 256):
    """Offsets of the per-rank slots of a staged placement (alac_b200_placement.slot_offsets) and the staging size:
    slot r starts on an `align` boundary and holds at least slot_bytes[r] (rank r's worst-case block)."""
    offs, at = [], 0
    for b in slot_bytes:
        offs.append(at)
        at += (int(b) + align - 1) // align * align
    return offs, at


def compaction_plan(totals, slot_offsets):
    """What the home rank does when every rank has reported its byte total: [(src offset in staging, dst offset in
    the job's buffer, bytes)] for ranks 1.. (rank 0 writes its block at offset 0 itself)."""
    out, at = [], int(totals[0])
    for r in range(1, len(totals)):
        if totals[r]:
            out.append((int(slot_offsets[r]), at, int(totals[r])))
        at += int(totals[r])
    return out, at


class SharedJob:
    """The single output buffer of a multi-rank encode job.

    capacity_bytes / total_packets size the packet buffer and the size table (worst case of the whole job).
    slot_bytes: per-rank worst-case block sizes (encode bounds) -> the staged form of alac_b200_placement: the home rank
    also owns a staging area with one reserved slot per rank, so packets travel while the ranks still compute; None
    selects the direct form (the assemble kernels store at the final offsets once every rank has been scanned).
    Every rank calls `placement(first_packet)` once per encode call (it advances the epoch in lock step)."""

    def __init__(self, engine, device, capacity_bytes: int, total_packets: int, home: int = 0, group=None, slot_bytes=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from .engine import EXCHANGE_BYTES, DevicePtr
        self.engine, self.device, self.home = engine, device, home
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.capacity, self.total_packets = int(capacity_bytes), int(total_packets)
        self.epoch = 0
        self.slot_offsets = None
        staging_bytes = 0
        if slot_bytes is not None:
            offs, staging_bytes = staging_slots(slot_bytes)
            self.slot_offsets = (C.c_uint64 * len(offs))(*offs)
        sizes_bytes = 4 * max(self.total_packets, 1)
        if self.rank == home:
            self._ptrs = [engine.device_alloc(self.capacity), engine.device_alloc(sizes_bytes), engine.device_alloc(EXCHANGE_BYTES),
                          engine.device_alloc(max(staging_bytes, 256) + 256)]       # (32 bytes of slack after the last slot, see alac_b200.h)
            tensor_from_pointer(self._ptrs[2], EXCHANGE_BYTES, device).zero_()
            torch.cuda.synchronize(device)
            handles = [engine.ipc_export(p) for p in self._ptrs]
        else:
            handles = None
        box = [handles]
        dist.broadcast_object_list(box, src=home, group=group)
        if self.rank != home:
            self._ptrs = [engine.ipc_open(h) for h in box[0]]
        self.packets_ptr, self.sizes_ptr, self.exchange_ptr, self.staging_ptr = self._ptrs
        # torch views exist only where the memory is local (the home rank)
        self.packets_region = DevicePtr(self.packets_ptr, self.capacity)
        self.packets = self.sizes = None
        if self.rank == home:
            self.packets = tensor_from_pointer(self.packets_ptr, self.capacity, device)
            self.sizes = tensor_from_pointer(self.sizes_ptr, max(self.total_packets, 1), device, "<i4")

    def placement(self, first_packet: int, defer_finish: bool = False):
        """defer_finish (staged form): the home rank's call returns once its own block is placed; call finish() when
        the job's buffer is needed (every rank's block at its final offset)."""
        from .engine import Placement
        self.epoch += 1
        staged = self.slot_offsets is not None
        return Placement(self.packets_ptr, self.capacity, self.sizes_ptr, int(first_packet), self.exchange_ptr,
                         self.rank, self.world, self.home, self.epoch, self.staging_ptr if staged else None,
                         self.slot_offsets if staged else None, 1 if (defer_finish and staged) else 0)

    def finish(self) -> int:
        """Ends a step whose placement asked for defer_finish: the home rank waits until the job's buffer is complete
        (returns the job's bytes), every other rank until its block has arrived on the home GPU.  A no-op otherwise."""
        return self.engine.placed_finish()

    def close(self):
        if getattr(self, "_ptrs", None) is None:
            return
        self.packets = self.sizes = None
        if self.rank == self.home:
            for p in self._ptrs:
                self.engine.device_free(p)
        else:
            for p in self._ptrs:
                self.engine.ipc_close(p)
        self._ptrs = None
