"""python scripts/multi_engine_check.py [n_devices]: ONE process, one engine over several GPUs
(alac_b200_engine_create_multi).  Encode and decode calls shard by frame range inside the C ABI; results are compared
byte for byte with a single-GPU engine, for device buffers (peer stores into the home GPU's buffer) and host buffers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import alac_b200
from tests import synth

n_dev = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
single = alac_b200.Engine(0)
multi = alac_b200.Engine(list(range(n_dev)))
assert multi.num_devices == n_dev
ok = True
for ch, depth, K, frames, streams in [(2, 16, 1, 4096 * 3000 + 11, None), (2, 24, 4, 4096 * 2000 + 999, None),
                                      (1, 32, 0, 4096 * 64, [(i * 4096 * 8, 4096 * 8 - (i % 3)) for i in range(8)]),
                                      (6, 24, 1, 4096 * 300 + 5, None)]:
    cfg = alac_b200.EncoderConfig(channels=ch, bit_depth=depth, frames_per_segment=K)
    pcm_d = synth.corpus_torch(0, frames, ch, depth, dev, seed=3)
    pcm_h = pcm_d.cpu().numpy()
    want = single.encode(pcm_d, cfg, streams=streams)
    got_d = multi.encode(pcm_d, cfg, streams=streams)
    got_h = multi.encode(pcm_h, cfg, streams=streams)
    same = (torch.equal(got_d.packets, want.packets) and torch.equal(torch.as_tensor(got_d.sizes), torch.as_tensor(want.sizes))
            and np.array_equal(got_h.packets, want.packets.cpu().numpy()) and np.array_equal(got_h.sizes.astype(np.int32), want.sizes.cpu().numpy()))
    dw = single.decode(want.cookie, want.packets, want.sizes)
    dd = multi.decode(want.cookie, want.packets, want.sizes)
    dh = multi.decode(want.cookie, got_h.packets, got_h.sizes)
    same_dec = torch.equal(dd.pcm, dw.pcm) and np.array_equal(dh.pcm, dw.pcm.cpu().numpy()) and torch.equal(dd.packet_samples, dw.packet_samples)
    print(f"{ch} ch {depth}-bit K={K}: encode {'OK' if same else 'MISMATCH'}, decode {'OK' if same_dec else 'MISMATCH'} "
          f"({want.num_packets} packets over {n_dev} GPUs)", flush=True)
    ok = ok and same and same_dec
print("multi_engine_check:", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
