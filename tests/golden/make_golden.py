"""Generates tests/golden/*.npz -- whole-packet golden vectors for the ALAC hot path.

Run in the build container (needs /root/reference): the packets are produced by the oracle's
frame drivers running on the REFERENCE's own unmodified dp_enc.c / ag_enc.c objects
(oracle/_ref/liboracle_ref.so), and every vector is checked to decode back through the reference's
dp_dec.c / ag_dec.c objects before it is written.  The reference itself ships no golden vectors
(SURVEY.md F7), so these pin the restated drivers + port primitives (tests/test_oracle.py) and the
CUDA path (tests/test_gpu_parity.py::test_golden_vectors) to the reference arithmetic.

    python tests/golden/make_golden.py
"""
import os
import sys
import wave

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def wav_head(name, frames):
    w = wave.open(f"/root/reference/audio/{name}.wav", "rb")
    ch, depth, sr = w.getnchannels(), w.getsampwidth() * 8, w.getframerate()
    w.setpos(44100 * 2)                     # a loud stretch of both fixtures
    pcm = np.frombuffer(w.readframes(frames), np.uint8).copy()
    return pcm, ch, depth, sr


CASES = [
    # name, maker, channels, depth, frames, K, fast
    ("wav05_mono16_k0", lambda: wav_head("05", 4096 * 3 + 100), None, None, None, 0, False),
    ("wav50_stereo16_k0", lambda: wav_head("50", 4096 * 3 + 100), None, None, None, 0, False),
    ("music_stereo16_k1", "music", 2, 16, 4096 * 2 + 500, 1, False),
    ("music_mono20_k2", "music", 1, 20, 4096 * 3 + 77, 2, False),
    ("music_stereo24_k0", "music", 2, 24, 4096 * 2 + 1904, 0, False),
    ("music_stereo32_k1", "music", 2, 32, 4096 + 300, 1, False),
    ("noise_stereo16_escape", "noise", 2, 16, 4096 + 64, 1, False),
    ("noise_mono24_escape", "noise", 1, 24, 4096 + 9, 1, False),
    ("silence_stereo16_runs", "silence", 2, 16, 4096 * 2, 1, False),
    ("music_8ch24_k1", "music", 8, 24, 4096 + 333, 1, False),
    ("music_5ch16_k0", "music", 5, 16, 4096 * 2, 0, False),
    ("music_stereo16_fast", "music", 2, 16, 4096 * 2 + 50, 1, True),
    ("tiny_tail_stereo16", "music", 2, 16, 4096 + 40, 1, False),
]


def main():
    assert O.have_reference(), "needs oracle/_ref (build it where /root/reference exists)"
    for name, maker, ch, depth, frames, K, fast in CASES:
        if callable(maker):
            pcm, ch, depth, sr = maker()
        else:
            sr = 44100
            pcm = synth.make(maker, frames, ch, depth, seed=len(name))
        enc = O.Encoder(ch, depth, sr, fast_mode=fast, reference=True)
        es = enc.encode_stream(pcm, K)
        back, st = O.Decoder(es.cookie, reference=True).decode_stream(es.packets, es.sizes)
        assert not st.any() and np.array_equal(back, pcm), name
        np.savez_compressed(os.path.join(HERE, name + ".npz"), pcm=pcm, packets=es.packets, sizes=es.sizes,
                            cookie=np.frombuffer(es.cookie, np.uint8), channels=ch, depth=depth, sample_rate=sr,
                            frames_per_segment=K, fast_mode=int(fast))
        print(f"{name:28s} ch={ch} depth={depth} pcm={pcm.nbytes} packets={es.packets.nbytes} n={len(es.sizes)}")


if __name__ == "__main__":
    main()
