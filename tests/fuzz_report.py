"""Developer aid: decode fuzzed packets on the GPU and with the oracle, report every packet both accept but decode differently."""
import glob, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # repo root
import numpy as np
import alac_b200
from oracle import oracle as O
O.build()
eng = alac_b200.Engine(0)
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
nbad = 0
nstat = 0
for seed in range(11, 17):
    rng = np.random.default_rng(seed)
    for ch, depth, golden in [(2, 16, "music_stereo16_k1"), (2, 24, "music_stereo24_k0"), (1, 20, "music_mono20_k2"),
                              (2, 32, "music_stereo32_k1"), (2, 16, "silence_stereo16_runs"), (2, 16, "noise_stereo16_escape")]:
        g = np.load([p for p in GOLD if golden in p][0])
        sizes = g["sizes"].astype(np.int64)
        offs = np.concatenate([[0], np.cumsum(sizes)])
        good = [g["packets"][offs[i]:offs[i + 1]].copy() for i in range(len(sizes))]
        pk, modes, src = [], [], []
        for rep in range(20):
            for i, p in enumerate(good):
                mode = int(rng.integers(0, 6))
                q = p.copy()
                if mode == 1:
                    for _ in range(int(rng.integers(1, 6))):
                        q[int(rng.integers(0, len(q)))] ^= 1 << int(rng.integers(0, 8))
                elif mode == 2:
                    q = q[: int(rng.integers(1, len(q)))]
                elif mode == 3:
                    q[int(rng.integers(0, min(8, len(q))))] = int(rng.integers(0, 256))
                elif mode == 4:
                    q = rng.integers(0, 256, int(rng.integers(1, 600)), dtype=np.uint8).astype(np.uint8)
                pk.append(q); modes.append(mode); src.append(i)
        blob, sz = np.concatenate(pk), np.array([len(q) for q in pk], np.uint32)
        cookie = bytes(g["cookie"])
        dec = eng.decode(cookie, blob, sz, raise_on_error=False)
        ref = O.Decoder(cookie)
        bpf = ch * (2 if depth == 16 else 4 if depth == 32 else 3)
        st, ns, pos = np.asarray(dec.packet_status), np.asarray(dec.packet_samples), 0
        for i, q in enumerate(pk):
            n = int(ns[i])
            # (bytes past the end of a short packet read as zero in the oracle and in the CUDA bit readers alike)
            want, rst = ref.decode_stream(q, np.array([len(q)], np.uint32))
            got = dec.pcm[pos * bpf:(pos + n) * bpf]
            pos += n
            if rst[0] == 0 and st[i] == 0 and not (len(want) == len(got) and np.array_equal(got, want)):
                nbad += 1
                p = good[src[i]]
                diff = np.nonzero(p != q)[0].tolist() if len(p) == len(q) else "len"
                first = int(np.nonzero(want[:min(len(want), len(got))] != got[:min(len(want), len(got))])[0][0]) if len(want) and len(got) else -1
                print(f"seed {seed} {golden} pkt {i} mode {modes[i]} src {src[i]} len {len(q)} n_gpu {n} n_ref {len(want)//bpf} "
                      f"diffbytes {diff if diff=='len' else [(d, hex(p[d]), hex(q[d])) for d in diff[:6]]} first_bad_byte {first} "
                      f"(frame {first // bpf if first >= 0 else -1}, byte-in-frame {first % bpf if first >= 0 else -1}) head {bytes(q[:14]).hex()}")
            elif (rst[0] == 0) != (st[i] == 0):
                nstat += 1
                if nstat <= 25: print(f"STATUS seed {seed} {golden} pkt {i} mode {modes[i]} src {src[i]} len {len(q)} gpu {st[i]} ref {rst[0]} n_gpu {n} head {bytes(q[:14]).hex()}")
print("bad", nbad, "status disagreements", nstat)
