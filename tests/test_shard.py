"""Multi-rank host logic on CPU: frame-range sharding plans and the rank-order packet gather
(gloo, world_size 2).  The oracle stands in for the GPU engine; the point is that sharding by
whole segments is byte-identical to the unsharded encode."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from alac_b200 import shard
from tests import synth


def test_plans_cover_and_align():
    for packets in [0, 1, 7, 100, 38760]:
        for world in [1, 2, 3, 8]:
            for K in [1, 2, 5, 64]:
                plan = shard.plan_packet_shards(packets, world, K)
                assert len(plan) == world
                at = 0
                for p0, n in plan:
                    assert p0 == at and (p0 % K == 0 or n == 0 or p0 == packets)
                    at += n
                assert at == packets
    assert shard.plan_packet_shards(10, 4, 0) == [(0, 10), (10, 0), (10, 0), (10, 0)]
    assert shard.plan_frame_shards(4096 * 3 + 5, 4096, 2, 1) == [(0, 8192), (8192, 4101)]


def _worker(rank, world, port, pcm, ch, depth, K, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    bpf = O.bytes_per_sample(depth) * ch
    a, n = shard.plan_frame_shards(pcm.nbytes // bpf, 4096, world, K)[rank]
    es = O.Encoder(ch, depth).encode_stream(pcm[a * bpf:(a + n) * bpf], K)
    pk, sz = shard.gather_packets(torch.from_numpy(es.packets), torch.from_numpy(es.sizes.astype(np.int32)))
    # the point-to-point form: only the destination rank (the last one here) receives, block by block
    res = shard.concat_packets_to(world - 1, torch.from_numpy(es.packets), torch.from_numpy(es.sizes.astype(np.int32)))
    assert (res is not None) == (rank == world - 1)
    if res is not None:
        assert torch.equal(res[0], pk) and torch.equal(res[1], sz)
    if rank == 0:
        q.put((pk.numpy(), sz.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("K", [1, 3])
def test_sharded_encode_is_byte_identical(oracle, K, world):
    ch, depth = 2, 16
    pcm = synth.make("music", 4096 * 7 + 321, ch, depth, seed=5)
    whole = oracle.Encoder(ch, depth).encode_stream(pcm, K)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + K + 10 * world
    procs = [ctx.Process(target=_worker, args=(r, world, port, pcm, ch, depth, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    pk, sz = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(sz.astype(np.uint32), whole.sizes)
    assert np.array_equal(pk, whole.packets)


# ---- staged placement: the host-side arithmetic of alac_b200_encode_placed's staged form, under gloo ------------------
def test_staging_slots_and_compaction_plan():
    offs, total = shard.staging_slots([1000, 0, 257, 4096])
    assert offs == [0, 1024, 1024, 1536] and total == 1536 + 4096
    plan, end = shard.compaction_plan([700, 0, 200, 50], offs)
    assert plan == [(1024, 700, 200), (1536, 900, 50)] and end == 950


def _staged_worker(rank, world, port, pcm, ch, depth, K, q):
    """Every rank encodes its frame range (oracle), writes its block into ITS SLOT of rank 0's staging area
    (send/recv stands in for the NVLink peer copy), reports its total; rank 0 compacts."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    bpf = O.bytes_per_sample(depth) * ch
    plan = shard.plan_frame_shards(pcm.nbytes // bpf, 4096, world, K)
    a, n = plan[rank]
    es = O.Encoder(ch, depth).encode_stream(pcm[a * bpf:(a + n) * bpf], K)
    bounds = [nn * bpf + ((nn + 4095) // 4096 + 1) * (7 * ch + 1) for _, nn in plan]
    offs, staging_bytes = shard.staging_slots(bounds)
    tot = torch.tensor([es.packets.nbytes], dtype=torch.int64)
    totals = [torch.zeros_like(tot) for _ in range(world)]
    dist.all_gather(totals, tot)                    # (the 1 KB exchange block on the device)
    totals = [int(t.item()) for t in totals]
    if rank != 0:
        dist.send(torch.from_numpy(es.packets.copy()), 0)
    else:
        staging = np.zeros(staging_bytes, np.uint8)
        out = np.zeros(sum(totals), np.uint8)
        out[:totals[0]] = es.packets                # the home rank owns offset 0
        for r in range(1, world):
            buf = torch.empty(totals[r], dtype=torch.uint8)
            dist.recv(buf, r)
            staging[offs[r]:offs[r] + totals[r]] = buf.numpy()
        moves, end = shard.compaction_plan(totals, offs)
        for src, dst, nb in moves:
            out[dst:dst + nb] = staging[src:src + nb]
        assert end == out.nbytes
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_staged_placement_is_byte_identical(oracle):
    ch, depth, K, world = 2, 24, 1, 2
    pcm = synth.make("music", 4096 * 5 + 77, ch, depth, seed=9)
    whole = oracle.Encoder(ch, depth).encode_stream(pcm, K)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_staged_worker, args=(r, world, port, pcm, ch, depth, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(out, whole.packets)
