"""torchrun --nproc-per-node N scripts/p2p_concat_check.py: every rank encodes its frame range on its own GPU, the
blocks are put in order on rank 0 with point-to-point copies (alac_b200.shard.concat_packets_to, NCCL send/recv),
and rank 0 compares with the unsharded encode and decodes the result."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import alac_b200
from alac_b200 import shard
from tests import synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ch, depth, F, K = 2, 24, 4096, 1
frames = F * 4001 + 777
cfg = alac_b200.EncoderConfig(channels=ch, bit_depth=depth, sample_rate=96000, frame_size=F, frames_per_segment=K)
bpf = cfg.bytes_per_frame
eng = alac_b200.Engine(local)
a, n = shard.plan_frame_shards(frames, F, world, K)[rank]
pcm = synth.corpus_torch(a, n, ch, depth, dev, seed=0)
enc = eng.encode(pcm, cfg)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
res = shard.concat_packets_to(0, enc.packets, torch.as_tensor(enc.sizes, device=dev).to(torch.int32))
ev1.record()
torch.cuda.synchronize()
if rank == 0:
    whole_pcm = synth.corpus_torch(0, frames, ch, depth, dev, seed=0)
    whole = eng.encode(whole_pcm, cfg)
    ok = torch.equal(res[0], whole.packets) and torch.equal(res[1].cpu(), torch.as_tensor(whole.sizes).cpu().to(torch.int32))
    dec = eng.decode(whole.cookie, res[0], res[1])
    ok = ok and dec.status == 0 and torch.equal(dec.pcm, whole_pcm)
    print(f"p2p concat over {world} ranks: {'OK' if ok else 'MISMATCH'}, {res[0].numel()} bytes in {ev0.elapsed_time(ev1):.2f} ms")
    assert ok
dist.barrier()
dist.destroy_process_group()
