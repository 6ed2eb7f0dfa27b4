"""torchrun --nproc-per-node N scripts/placed_check.py [depth]: the cross-GPU step inside the product.

Every rank encodes its frame range with alac_b200_encode_placed: the packet-offset exchange runs on the devices (a 1 KB
block in rank 0's memory) and every rank's assemble kernel stores its packets at their final offset in ONE buffer on
rank 0, over NVLink (CUDA IPC mapping, no NCCL on the data path).  Checked: the buffer equals the unsharded encode byte
for byte (packets and size table); every rank then decodes its packet range straight out of that buffer (peer loads)
and gets its PCM back."""
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
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ch, F, K = 2, 4096, 1
frames = F * 6001 + 777
cfg = alac_b200.EncoderConfig(channels=ch, bit_depth=depth, sample_rate=96000, frame_size=F, frames_per_segment=K)
eng = alac_b200.Engine(local)
plan = shard.plan_frame_shards(frames, F, world, K)
pk_plan = shard.plan_packet_shards((frames + F - 1) // F, world, K)
a, n = plan[rank]
pcm = synth.corpus_torch(a, n, ch, depth, dev, seed=0)
form = sys.argv[2] if len(sys.argv) > 2 else "staged"
slots = [alac_b200.encode_bound(cfg, nn) for _, nn in plan] if form == "staged" else None
job = shard.SharedJob(eng, dev, alac_b200.encode_bound(cfg, frames, world), (frames + F - 1) // F, slot_bytes=slots)
ok = True
for it in range(3):                                  # several epochs through the same exchange block
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    sizes, npk, nb, base, mine, stats = eng.encode_placed(pcm, cfg, job.placement(pk_plan[rank][0], defer_finish=(it == 1)))
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    # decode this rank's block: its own copy, and straight out of the shared buffer (peer loads)
    dec = eng.decode(alac_b200.magic_cookie(cfg), mine, sizes)
    ok = ok and dec.status == 0 and torch.equal(dec.pcm, pcm)
    job.finish()
    base = eng.placed_base()        # (a deferred call of a rank other than the home rank learns its offset at finish())
    dist.barrier()          # (the shared buffer is complete once the home rank's call -- or its finish() -- has returned)
    dec = eng.decode(alac_b200.magic_cookie(cfg), job.packets_region[base:base + nb], sizes)
    ok = ok and dec.status == 0 and torch.equal(dec.pcm, pcm)
    tot = torch.tensor([nb], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    if rank == 0:
        whole_pcm = synth.corpus_torch(0, frames, ch, depth, dev, seed=0)
        whole = eng.encode(whole_pcm, cfg)
        same = int(tot.item()) == whole.nbytes and torch.equal(job.packets[:whole.nbytes], whole.packets) and \
            torch.equal(job.sizes[:whole.num_packets], torch.as_tensor(whole.sizes, device=dev).to(torch.int32))
        ok = ok and same
        print(f"[{it}] {form} placed encode over {world} ranks, {depth}-bit: {'OK' if same else 'MISMATCH'}; {int(tot.item())} bytes, rank 0 call {ms:.2f} ms "
              f"(kernels {stats['ms_kernels']:.2f} ms)", flush=True)
        job.packets[:whole.nbytes].zero_()
    dist.barrier()
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("placed_check:", "OK" if int(flag.item()) else "FAILED")
job.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
