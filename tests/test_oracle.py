"""CPU tests of the oracle (test infrastructure): pins the restated primitives and drivers to the
reference's arithmetic.  No GPU needed.

  * SURVEY.md Appendix D known-answer vectors (made from the reference's own objects)
  * restated primitives == reference objects (oracle/_ref) on random inputs, when _ref is present
  * golden whole-packet vectors in tests/golden/ (made through the reference's primitives)
  * closed-form anchors: cookie bytes, init_coefs, the 26-bit silence stream
"""
import glob
import os

import numpy as np
import pytest

from tests import synth

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _kat_input(kind, n):
    s = 12345
    x = np.zeros(n, np.int64)
    for i in range(n):
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        tri = abs((i % 128) - 64) - 32
        if kind == "tri":
            x[i] = (s >> 22) - 512 + 125 * tri
        elif kind == "noise16":
            x[i] = (s >> 16) - 32768
        elif kind == "bursts":
            x[i] = 0 if (i % 1000 < 900) else (s >> 26) - 32
    return x.astype(np.int32)


# SURVEY.md Appendix D
KATS = [
    ("tri", 4096, 8, 16, "c64bab05", "4ff0743c", [1209, -756, -3, 15, 54, 33, -9, -61], 49328, "9a3f8805"),
    ("tri", 4096, 4, 16, "c64bab05", "ba1dc554", [656, -179, 54, 17], 47717, "db4eefe8"),
    ("tri", 4096, 8, 17, "c64bab05", "4ff0743c", [1209, -756, -3, 15, 54, 33, -9, -61], 49329, "83dc7499"),
    ("tri", 512, 8, 17, "d7b43b59", "a3e2a6e5", [1250, -868, -19, 74, 61, 58, 35, 37], 6249, "bbcb8b1a"),
    ("tri", 1904, 8, 16, "596fe3f1", "8d5c118c", [1215, -826, -19, 55, 54, 39, 15, 2], 22996, "aaacbb79"),
    ("silence", 4096, 8, 16, "38699dc5", "38699dc5", [1216, -928, -64, 0, 0, 0, 0, 0], 26, "b3ed6308"),
    ("noise16", 4096, 8, 17, "895224a6", "933962bc", [1189, -1175, -96, -87, -59, -51, -86, -145], 101868, "34d234ef"),
    ("bursts", 4096, 4, 16, "87f80d24", "d1a508e6", [1115, -789, -46, 52], 4470, "b8616c42"),
]


@pytest.mark.parametrize("reference", [False, True])
@pytest.mark.parametrize("kat", KATS, ids=[f"{k[0]}-{k[1]}-{k[2]}-{k[3]}" for k in KATS])
def test_primitive_kats(oracle, kat, reference):
    if reference and not oracle.have_reference():
        pytest.skip("oracle/_ref not built")
    kind, n, taps, cb, fx, fpc, coefs, bits, fb = kat
    x = _kat_input(kind, n)
    c = oracle.init_coefs()
    pc = oracle.pc_block(x, c, taps, cb, reference=reference)
    data, nb, st = oracle.dyn_comp(pc, cb, reference=reference)
    assert st == 0
    assert "%08x" % oracle.fnv1a(x) == fx
    assert "%08x" % oracle.fnv1a(pc) == fpc
    assert list(c[:taps]) == coefs
    assert nb == bits
    assert "%08x" % oracle.fnv1a(data[:(nb + 7) // 8]) == fb
    r, nb2, st2 = oracle.dyn_decomp(data, n, cb, reference=reference)
    assert st2 == 0 and nb2 == nb and np.array_equal(r, pc)
    c2 = oracle.init_coefs()
    y = oracle.unpc_block(r, c2, taps, cb, reference=reference)
    assert np.array_equal(y, x) and np.array_equal(c2, c)


def test_closed_form_anchors(oracle):
    assert list(oracle.init_coefs()[:4]) == [1216, -928, -64, 0]
    data, nb, _ = oracle.dyn_comp(np.zeros(4096, np.int32), 16)
    assert nb == 26 and bytes(data[:4]) == bytes.fromhex("7fc3ffc0")
    ck = oracle.Encoder(2, 16, 44100).cookie()
    assert ck.hex() == "000010000010280a0e0200ff00000000000000000000ac44"
    ck8 = oracle.Encoder(8, 24, 48000).cookie()
    assert len(ck8) == 48 and ck8[24:36] == bytes([0, 0, 0, 24]) + b"chan" + bytes(4)
    assert ck8[36:40] == ((127 << 16) | 8).to_bytes(4, "little")      # layout tag stored native-endian


@pytest.mark.parametrize("taps", [4, 8, 5, 16, 31, 0])
@pytest.mark.parametrize("chanbits", [16, 17, 21, 24])
def test_predictor_port_equals_reference(oracle, taps, chanbits):
    if not oracle.have_reference():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(taps * 100 + chanbits)
    for n in [1, 3, taps + 1, taps + 2, 100, 1000]:
        n = max(n, 1)
        amp = 1 << (chanbits - 2)
        x = (rng.integers(-amp, amp, n) // rng.integers(1, 64)).astype(np.int32)
        ca, cb_ = oracle.init_coefs(32), oracle.init_coefs(32)
        ra = oracle.pc_block(x, ca, taps, chanbits)
        rb = oracle.pc_block(x, cb_, taps, chanbits, reference=True)
        m = max(n, 1)
        assert np.array_equal(ra[:m], rb[:m]) and np.array_equal(ca, cb_)
        ca, cb_ = oracle.init_coefs(32), oracle.init_coefs(32)
        ya = oracle.unpc_block(ra, ca, taps, chanbits)
        yb = oracle.unpc_block(ra, cb_, taps, chanbits, reference=True)
        assert np.array_equal(ya, yb) and np.array_equal(ca, cb_)


@pytest.mark.parametrize("bit_size", [16, 17, 21, 24, 32])
def test_golomb_port_equals_reference(oracle, bit_size):
    if not oracle.have_reference():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(bit_size)
    lim = 1 << (min(bit_size, 31) - 1)
    for scale in [1, 4, 50, 3000, lim]:
        for n in [1, 2, 17, 512, 4096]:
            x = rng.integers(-scale, scale + 1, n).astype(np.int64)
            x[rng.random(n) < 0.4] = 0                      # zero runs
            x = np.clip(x, -lim, lim - 1).astype(np.int32)
            for start in (0, 3):
                a, na, sa = oracle.dyn_comp(x, bit_size, start_bit=start)
                b, nb_, sb = oracle.dyn_comp(x, bit_size, reference=True, start_bit=start)
                assert (na, sa) == (nb_, sb)
                mask0 = 0xFF >> start
                assert (a[0] & mask0) == (b[0] & mask0) and np.array_equal(a[1:], b[1:])
                ra, ma, ta = oracle.dyn_decomp(a, n, bit_size, start_bit=start)
                rb, mb, tb = oracle.dyn_decomp(a, n, bit_size, reference=True, start_bit=start)
                assert (ma, ta) == (mb, tb) == (na, 0)
                assert np.array_equal(ra, x) and np.array_equal(rb, x)


def test_golomb_decoder_overrun_is_param_error(oracle):
    x = (np.arange(512) * 37 % 2000 - 1000).astype(np.int32)
    data, nb, _ = oracle.dyn_comp(x, 16)
    for ref in ([False, True] if oracle.have_reference() else [False]):
        _, _, st = oracle.dyn_decomp(data[: len(data) // 2], 512, 16, reference=ref, cap_bytes=len(data) // 2)
        assert st == -50


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_packets(oracle, path):
    """The restated (port) primitives under the restated drivers reproduce the golden packets that
    were made with the reference's own primitives; decode gives the PCM back."""
    g = np.load(path)
    ch, depth, sr, K, fast = int(g["channels"]), int(g["depth"]), int(g["sample_rate"]), int(g["frames_per_segment"]), bool(g["fast_mode"])
    enc = oracle.Encoder(ch, depth, sr, fast_mode=fast)
    es = enc.encode_stream(g["pcm"], K)
    assert es.cookie == bytes(g["cookie"])
    assert np.array_equal(es.sizes, g["sizes"])
    assert np.array_equal(es.packets, g["packets"])
    back, st = oracle.Decoder(es.cookie).decode_stream(g["packets"], g["sizes"])
    assert not st.any() and np.array_equal(back, g["pcm"])


def test_golden_present():
    assert len(GOLDEN) >= 10


def test_decoder_error_paths(oracle):
    g = np.load(GOLDEN[0])
    dec = oracle.Decoder(bytes(g["cookie"]))
    first = g["packets"][: int(g["sizes"][0])].copy()
    _, _, st = dec.decode_packet(first)
    assert st == 0
    _, _, st = dec.decode_packet(first[: len(first) // 2])           # truncated
    assert st == -50
    bad = first.copy()
    bad[1] |= 0x10                                                   # unused header bits must be zero
    _, _, st = dec.decode_packet(bad)
    assert st == -50


def _bits(data):
    return "".join(f"{b:08b}" for b in data)


def _pack(bits):
    bits = bits + "0" * (-len(bits) % 8)
    return np.frombuffer(int(bits, 2).to_bytes(len(bits) // 8, "big"), np.uint8).copy()


def test_decoder_skips_fil_and_dse(oracle):
    """FIL / DSE elements in front of the audio element are parsed and ignored
    (codec/ALACDecoder.cu:941-953, :1012-1059)."""
    path = [p for p in GOLDEN if "music_stereo16_k1" in p][0]
    g = np.load(path)
    dec = oracle.Decoder(bytes(g["cookie"]))
    first = g["packets"][: int(g["sizes"][0])]
    want, n, st = dec.decode_packet(first)
    assert st == 0
    fil = "110" + "0011" + "10101010" * 3                            # ID_FIL, count 3, 3 bytes
    dse = "100" + "0000" + "1" + "00000010"                          # ID_DSE, tag 0, align, count 2
    pre = fil + dse
    pre += "0" * (-len(pre) % 8) + "11110000" * 2                    # byte-align then 2 data bytes
    got, n2, st = dec.decode_packet(_pack(pre + _bits(first)))
    assert st == 0 and n2 == n and np.array_equal(got, want)


@pytest.mark.parametrize("ch,depth", [(1, 16), (2, 16), (2, 20), (1, 24), (2, 24), (2, 32), (3, 16), (6, 24), (8, 24)])
def test_round_trip_identity(oracle, ch, depth):
    for kind in ["music", "noise", "silence", "square"]:
        pcm = synth.make(kind, 4096 + 1000, ch, depth, seed=3)
        es = oracle.Encoder(ch, depth).encode_stream(pcm, 0)
        back, st = oracle.Decoder(es.cookie).decode_stream(es.packets, es.sizes)
        assert not st.any() and np.array_equal(back, pcm)


def _one_walk_pc_block(x, coefs, taps, chanbits, denshift=9):
    """The kernels' formulation of pc_block's 4 / 8-tap paths (alac_device.cuh lms_adapt / predict_enc_step), restated in
    Python: both error signs are ONE walk on left = |err| with left -= w * ((|b| + c) >> denshift), c = 0 or 2^denshift - 1,
    a tap is updated while left > 0, and the update adds sign(b) * (+1 for err < 0, -1 otherwise)."""
    def sext(v, bits):
        v &= (1 << bits) - 1
        return v - (1 << bits) if v >> (bits - 1) else v
    a = [int(c) for c in coefs[:taps]]
    n = len(x)
    res = [0] * n
    res[0] = int(x[0])
    for j in range(1, min(taps + 1, n)):
        res[j] = sext(int(x[j]) - int(x[j - 1]), chanbits)
    for j in range(taps + 1, n):
        top = int(x[j - taps - 1])
        b = [top - int(x[j - 1 - k]) for k in range(taps)]
        acc = sext((1 << (denshift - 1)) - sum(a[k] * b[k] for k in range(taps)), 32)     # (int32 arithmetic, as in C and on the GPU)
        err = sext(int(x[j]) - top - (acc >> denshift), chanbits)
        res[j] = err
        m = -1 if err < 0 else 0
        nsg, c, left = -2 * m - 1, m & ((1 << denshift) - 1), abs(err)
        for k in range(taps - 1, -1, -1):
            sb = (b[k] > 0) - (b[k] < 0)
            if left > 0:
                a[k] = sext(a[k] + sb * nsg, 16)
            left -= (taps - k) * ((sb * b[k] + c) >> denshift)
    coefs[:taps] = np.array(a, np.int16)
    return np.array(res, np.int32)


@pytest.mark.parametrize("taps", [4, 8])
@pytest.mark.parametrize("chanbits", [16, 17, 21, 25])
def test_one_walk_ladder_equals_the_reference_pc_block(oracle, taps, chanbits):
    """The sign-LMS ladder as the CUDA kernels compute it (one walk on |err|, predicated update: DESIGN.md section 4) against
    the reference's own pc_block object (codec/dp_enc.c:236-329), residuals and adapted coefficients, on signals that make
    the walk stop early, run through, hit err = 0 and hit history differences of 0 and +-1."""
    use_ref = oracle.have_reference()
    rng = np.random.default_rng(taps * 31 + chanbits)
    amp = 1 << (chanbits - 2)
    signals = [
        (rng.integers(-amp, amp, 1500) // rng.integers(1, 64)).astype(np.int32),                        # noise of one scale
        (np.cumsum(rng.integers(-3, 4, 1500)) + (amp // 4 * np.sin(np.arange(1500) / 17.0))).astype(np.int32),   # smooth: small errors, walk stops early
        np.repeat(rng.integers(-amp, amp, 150), 10).astype(np.int32),                                   # plateaus: b = 0, err = 0
        (rng.integers(-1, 2, 1500)).astype(np.int32),                                                   # |b| <= 2
    ]
    for x in signals:
        ca, cb_ = oracle.init_coefs(32), oracle.init_coefs(32)
        want = oracle.pc_block(x, ca, taps, chanbits, reference=use_ref)
        got = _one_walk_pc_block(x, cb_, taps, chanbits)
        assert np.array_equal(got, want[:len(x)])
        assert np.array_equal(ca[:taps], cb_[:taps])
