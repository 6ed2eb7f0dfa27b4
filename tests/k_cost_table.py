"""Compression-ratio cost of the encoder-reset schedule: the oracle encodes the
first minutes of the two bench corpora with a fresh encoder every K frames (K = 0: never, what alacconvert does).
    python tests/k_cost_table.py [seconds]       -> JSON rows (test infrastructure: uses oracle/)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import oracle as O
from tests import synth

secs = int(sys.argv[1]) if len(sys.argv) > 1 else 120
rows = []
for name, depth, rate in (("c2 16-bit/44.1 kHz stereo", 16, 44100), ("c3 24-bit/96 kHz stereo", 24, 96000)):
    frames = secs * rate
    pcm = synth.corpus_torch(0, frames, 2, depth, torch.device("cpu")).numpy()
    row = {"corpus": name, "seconds": secs, "packets": (frames + 4095) // 4096, "pcm_bytes": int(pcm.size), "ratio": {}}
    for K in (0, 64, 8, 1):
        enc = O.Encoder(2, depth, rate).encode_stream(pcm, K)
        row["ratio"]["K=%d" % K] = round(float(enc.packets.size) / pcm.size, 5)
    rows.append(row)
    print(json.dumps(row), flush=True)
