"""Developer probe: host-buffer (pinned) encode/decode phase times."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import alac_b200
from tests import synth
secs = int(sys.argv[1]) if len(sys.argv) > 1 else 3600
dev = torch.device("cuda", 0)
frames = secs * 44100
pcm_d = torch.cat([synth.corpus_torch(a, min(1 << 24, frames - a), 2, 16, dev) for a in range(0, frames, 1 << 24)])
cfg = alac_b200.EncoderConfig(channels=2, bit_depth=16, frames_per_segment=1)
eng = alac_b200.Engine(0)
pcm_h = torch.empty(pcm_d.numel(), dtype=torch.uint8).pin_memory(); pcm_h.copy_(pcm_d)
bound = alac_b200.encode_bound(cfg, frames); npk = (frames + 4095) // 4096
pk_h = torch.empty(bound, dtype=torch.uint8).pin_memory(); sz_h = torch.empty(npk, dtype=torch.int32).pin_memory()
out_h = torch.empty(pcm_d.numel(), dtype=torch.uint8).pin_memory()
pcm_np, pk_np, sz_np, out_np = pcm_h.numpy(), pk_h.numpy(), sz_h.numpy().view(np.uint32), out_h.numpy()
for i in range(4):
    t0 = time.perf_counter(); e = eng.encode(pcm_np, cfg, out=pk_np, out_sizes=sz_np); t1 = time.perf_counter()
    d = eng.decode(e.cookie, e.packets, e.sizes, out=out_np); t2 = time.perf_counter()
    print(json.dumps({"enc_wall_ms": round((t1 - t0) * 1e3, 2), "dec_wall_ms": round((t2 - t1) * 1e3, 2),
                      "enc": {k: round(v, 2) for k, v in e.stats.items() if k.startswith("ms_")},
                      "dec": {k: round(v, 2) for k, v in d.stats.items() if k.startswith("ms_")}}))
assert np.array_equal(d.pcm, pcm_np)
# raw PCIe reference: one big pinned copy each way
t = torch.empty_like(pcm_d)
torch.cuda.synchronize(); t0 = time.perf_counter(); t.copy_(pcm_h, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
out_h.copy_(t, non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter()
print("raw H2D GB/s", pcm_d.numel() / (t1 - t0) / 1e9, "raw D2H GB/s", pcm_d.numel() / (t2 - t1) / 1e9)
