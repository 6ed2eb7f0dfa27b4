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
