"""Developer micro-benchmark: encode/decode kernel times on the bench corpus (device resident)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import alac_b200
from tests import synth

secs = int(sys.argv[1]) if len(sys.argv) > 1 else 3600
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ch = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rate = 44100
dev = torch.device("cuda", 0)
frames = secs * rate
pcm = torch.cat([synth.corpus_torch(a, min(1 << 24, frames - a), ch, depth, dev) for a in range(0, frames, 1 << 24)])
cfg = alac_b200.EncoderConfig(channels=ch, bit_depth=depth, sample_rate=rate, frames_per_segment=1)
eng = alac_b200.Engine(0)
enc = eng.encode(pcm, cfg)
out = torch.empty_like(pcm)
res = {}
for _ in range(3):
    e = eng.encode(pcm, cfg)
    d = eng.decode(e.cookie, e.packets, e.sizes, out=out)
    res = {"lib": os.path.basename(os.environ.get("ALAC_B200_LIB", "default")), "depth": depth, "channels": ch, "sample_frames": frames,
           "search_ms": round(e.stats["ms_search"], 3), "final_ms": round(e.stats["ms_final"], 3),
           "asm_ms": round(e.stats["ms_assemble"], 3), "enc_ms": round(e.stats["ms_kernels"], 3),
           "entropy_ms": round(d.stats["ms_entropy"], 3), "finish_ms": round(d.stats["ms_finish"], 3), "dec_all_ms": round(d.stats["ms_kernels"], 3),
           "ratio": round(e.nbytes / pcm.numel(), 4),
           "roundtrip_msamples_s": round(frames / ((e.stats["ms_kernels"] + d.stats["ms_kernels"]) / 1e3) / 1e6, 1)}
assert torch.equal(d.pcm, pcm)
print(json.dumps(res))
