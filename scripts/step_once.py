"""One encode + decode step of a bench corpus for profiling (no CPU baseline, no host copies):
    python scripts/step_once.py <seconds> <depth> <rate> [warmup]
Warm-up steps run outside the cudaProfilerStart/Stop window, so `ncu --profile-from-start off` sees exactly one step."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import alac_b200
from tests import synth

secs = int(sys.argv[1]) if len(sys.argv) > 1 else 3600
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rate = int(sys.argv[3]) if len(sys.argv) > 3 else 44100
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = torch.device("cuda", 0)
frames = secs * rate
pcm = torch.cat([synth.corpus_torch(a, min(1 << 24, frames - a), 2, depth, dev) for a in range(0, frames, 1 << 24)])
cfg = alac_b200.EncoderConfig(channels=2, bit_depth=depth, sample_rate=rate, frames_per_segment=1)
eng = alac_b200.Engine(0)
out = torch.empty_like(pcm)
for _ in range(warm):
    e = eng.encode(pcm, cfg)
    d = eng.decode(e.cookie, e.packets, e.sizes, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
e = eng.encode(pcm, cfg)
d = eng.decode(e.cookie, e.packets, e.sizes, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
assert torch.equal(d.pcm, pcm)
print(json.dumps({"depth": depth, "rate": rate, "seconds": secs, "packets": int(e.num_packets), "bytes": int(e.nbytes),
                  "search_ms": round(e.stats["ms_search"], 3), "final_ms": round(e.stats["ms_final"], 3),
                  "asm_ms": round(e.stats["ms_assemble"], 3), "enc_ms": round(e.stats["ms_kernels"], 3),
                  "entropy_ms": round(d.stats["ms_entropy"], 3), "finish_ms": round(d.stats["ms_finish"], 3),
                  "dec_ms": round(d.stats["ms_kernels"], 3)}))
