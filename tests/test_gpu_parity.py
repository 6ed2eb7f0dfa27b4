"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit-exact.

Packets and cookie on encode, PCM on decode, and round-trip identity; covers every depth,
mono / stereo / multichannel, partial and tiny tail frames, escape and zero-run content,
encoder-reset schedules K = 1, K > 1 and K = 0 (one serial chain per stream).
"""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu


def _oracle_encode(O, pcm, ch, depth, K, frame_size=4096, fast=False, sr=44100):
    enc = O.Encoder(ch, depth, sr, frame_size=frame_size, fast_mode=fast, reference=O.have_reference())
    return enc.encode_stream(pcm, K)


def _check_encode(engine, O, pcm, ch, depth, K, frame_size=4096, fast=False):
    import alac_b200
    cfg = alac_b200.EncoderConfig(channels=ch, bit_depth=depth, sample_rate=44100, frame_size=frame_size,
                                  fast_mode=fast, frames_per_segment=K)
    got = engine.encode(pcm, cfg)
    want = _oracle_encode(O, pcm, ch, depth, K, frame_size, fast)
    assert got.cookie == want.cookie
    assert got.num_packets == len(want.sizes)
    assert np.array_equal(np.asarray(got.sizes, np.uint32), want.sizes), \
        f"first size mismatch at packet {int(np.argmax(np.asarray(got.sizes) != want.sizes))}"
    assert got.nbytes == want.packets.nbytes
    if not np.array_equal(got.packets, want.packets):
        bad = int(np.argmax(got.packets != want.packets))
        offs = np.concatenate([[0], np.cumsum(want.sizes)])
        pk = int(np.searchsorted(offs, bad, side="right") - 1)
        raise AssertionError(f"packet bytes differ: byte {bad} (packet {pk}, byte {bad - offs[pk]} of {want.sizes[pk]})")
    return got, want


def _check_decode(engine, O, enc, pcm):
    dec = engine.decode(enc.cookie, enc.packets, np.asarray(enc.sizes, np.uint32))
    assert dec.status == 0
    want, st = O.Decoder(enc.cookie, reference=O.have_reference()).decode_stream(np.asarray(enc.packets), np.asarray(enc.sizes, np.uint32))
    assert not st.any()
    assert np.array_equal(dec.pcm, want)
    assert np.array_equal(dec.pcm, pcm), "round trip is not the identity"


@pytest.mark.parametrize("depth", [16, 20, 24, 32])
@pytest.mark.parametrize("ch", [1, 2])
@pytest.mark.parametrize("kind", ["music", "noise", "silence"])
def test_encode_decode_k1(engine, oracle, depth, ch, kind):
    frames = 4096 * 5 + 1904
    pcm = synth.make(kind, frames, ch, depth, seed=depth + ch)
    got, _ = _check_encode(engine, oracle, pcm, ch, depth, K=1)
    _check_decode(engine, oracle, got, pcm)


@pytest.mark.parametrize("K", [0, 3])
@pytest.mark.parametrize("ch,depth", [(2, 16), (1, 24), (2, 24)])
def test_encode_chained_segments(engine, oracle, K, ch, depth):
    pcm = synth.make("music", 4096 * 7 + 100, ch, depth, seed=9)
    got, _ = _check_encode(engine, oracle, pcm, ch, depth, K=K)
    _check_decode(engine, oracle, got, pcm)


@pytest.mark.parametrize("ch,depth", [(3, 16), (6, 24), (8, 24), (5, 20), (4, 32), (7, 16)])
def test_multichannel(engine, oracle, ch, depth):
    pcm = synth.make("music", 4096 * 2 + 777, ch, depth, seed=ch)
    got, _ = _check_encode(engine, oracle, pcm, ch, depth, K=1)
    _check_decode(engine, oracle, got, pcm)


@pytest.mark.parametrize("tail", [1, 2, 7, 8, 9, 31, 32, 33, 71, 72, 100, 287, 288, 4095])
@pytest.mark.parametrize("ch,depth", [(2, 16), (1, 16), (2, 24)])
def test_tiny_tails(engine, oracle, tail, ch, depth):
    pcm = synth.make("music", 4096 + tail, ch, depth, seed=tail)
    got, _ = _check_encode(engine, oracle, pcm, ch, depth, K=1)
    _check_decode(engine, oracle, got, pcm)


def test_square_full_scale(engine, oracle):
    for ch, depth in [(2, 16), (1, 16), (2, 24), (2, 32)]:
        pcm = synth.make("square", 4096 * 3, ch, depth)
        got, _ = _check_encode(engine, oracle, pcm, ch, depth, K=1)
        _check_decode(engine, oracle, got, pcm)


def test_fast_mode(engine, oracle):
    for kind in ["music", "noise"]:
        pcm = synth.make(kind, 4096 * 3 + 50, 2, 16, seed=4)
        got, _ = _check_encode(engine, oracle, pcm, 2, 16, K=1, fast=True)
        _check_decode(engine, oracle, got, pcm)


def test_small_frame_size(engine, oracle):
    pcm = synth.make("music", 1024 * 9 + 300, 2, 16, seed=5)
    got, _ = _check_encode(engine, oracle, pcm, 2, 16, K=2, frame_size=1024)
    _check_decode(engine, oracle, got, pcm)


def test_multi_stream_batch(engine, oracle):
    import alac_b200
    ch, depth = 2, 16
    lens = [4096 * 2 + 5, 4096, 300, 4096 * 3]
    parts = [synth.make("music", n, ch, depth, seed=10 + i) for i, n in enumerate(lens)]
    pcm = np.concatenate(parts)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    cfg = alac_b200.EncoderConfig(channels=ch, bit_depth=depth, frames_per_segment=0)
    got = engine.encode(pcm, cfg, streams=[(int(a), int(n)) for a, n in zip(starts, lens)])
    want_p, want_s = [], []
    for part in parts:
        es = _oracle_encode(oracle, part, ch, depth, 0)
        want_p.append(es.packets)
        want_s.append(es.sizes)
    assert np.array_equal(np.asarray(got.sizes, np.uint32), np.concatenate(want_s))
    assert np.array_equal(got.packets, np.concatenate(want_p))


def test_device_buffers(engine, oracle):
    """Same call with device-resident input/output (torch CUDA tensors)."""
    import torch
    import alac_b200
    pcm = synth.make("music", 4096 * 4 + 123, 2, 16, seed=21)
    cfg = alac_b200.EncoderConfig(channels=2, bit_depth=16, frames_per_segment=1)
    got = engine.encode(torch.from_numpy(pcm).cuda(), cfg)
    want = _oracle_encode(oracle, pcm, 2, 16, 1)
    assert np.array_equal(got.packets.cpu().numpy(), want.packets)
    assert np.array_equal(got.sizes.cpu().numpy().astype(np.uint32), want.sizes)
    dec = engine.decode(got.cookie, got.packets, got.sizes)
    assert np.array_equal(dec.pcm.cpu().numpy(), pcm)


def test_coef_state_streaming(engine, oracle):
    """K = 0 fed in two calls with the coefficient state carried equals one call (and the oracle)."""
    import alac_b200
    pcm = synth.make("music", 4096 * 6, 2, 16, seed=33)
    cfg = alac_b200.EncoderConfig(channels=2, bit_depth=16, frames_per_segment=0)
    half = 4096 * 3 * cfg.bytes_per_frame
    state = np.zeros((1, 256), np.int16)
    # a fresh state equals init_coefs rows
    init = np.zeros(8, np.int16)
    init[:3] = [1216, -928, -64]
    state[:] = np.tile(init, 32)
    a = engine.encode(pcm[:half], cfg, coef_state=state)
    b = engine.encode(pcm[half:], cfg, coef_state=state)
    want = _oracle_encode(oracle, pcm, 2, 16, 0)
    assert np.array_equal(np.concatenate([a.packets, b.packets]), want.packets)


# ---------------------------------------------------------------------------------------------
# golden vectors (made through the reference's own primitives, tests/golden/make_golden.py)
# ---------------------------------------------------------------------------------------------
import glob
import os
import subprocess

_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("path", _GOLDEN, ids=[os.path.basename(p)[:-4] for p in _GOLDEN])
def test_golden_vectors(engine, path):
    import alac_b200
    g = np.load(path)
    cfg = alac_b200.EncoderConfig(channels=int(g["channels"]), bit_depth=int(g["depth"]), sample_rate=int(g["sample_rate"]),
                                  fast_mode=bool(g["fast_mode"]), frames_per_segment=int(g["frames_per_segment"]))
    got = engine.encode(g["pcm"], cfg)
    assert got.cookie == bytes(g["cookie"])
    assert np.array_equal(np.asarray(got.sizes, np.uint32), g["sizes"])
    assert np.array_equal(got.packets, g["packets"])
    dec = engine.decode(bytes(g["cookie"]), g["packets"], g["sizes"])
    assert dec.status == 0 and np.array_equal(dec.pcm, g["pcm"])


def test_decode_error_status(engine):
    """Truncated / corrupt packets give kALAC_ParamError per packet; good packets still decode."""
    import alac_b200
    g = np.load([p for p in _GOLDEN if "music_stereo16_k1" in p][0])
    sizes = g["sizes"].copy()
    packets = g["packets"].copy()
    bad = packets.copy()
    bad[1] |= 0x10                                  # unused header bits of packet 0 must be zero
    dec = engine.decode(bytes(g["cookie"]), bad, sizes, raise_on_error=False)
    assert dec.status == -50
    assert int(dec.packet_status[0]) == -50 and not np.any(np.asarray(dec.packet_status[1:]))
    # truncate the stream in the middle of packet 0
    cut = int(sizes[0]) // 2
    dec = engine.decode(bytes(g["cookie"]), packets[:cut].copy(), np.array([cut], np.uint32), raise_on_error=False)
    assert dec.status == -50


def test_decode_mixed_short_packets(engine, oracle):
    """Config-5 shape at test size: many short packets with mixed frame sizes (32-bit samples)."""
    rng = np.random.default_rng(0)
    ch, depth = 2, 32
    enc = oracle.Encoder(ch, depth, 48000, reference=oracle.have_reference())
    pk, sz, pcms = [], [], []
    for i in range(300):
        n = int(rng.integers(32, 4097))
        pcm = synth.make("music", n, ch, depth, seed=i)
        enc.reset()
        p = enc.encode_packet(pcm, n)
        pk.append(p); sz.append(len(p)); pcms.append(pcm)
    packets, sizes = np.concatenate(pk), np.array(sz, np.uint32)
    dec = engine.decode(enc.cookie(), packets, sizes)
    assert dec.status == 0
    assert np.array_equal(dec.pcm, np.concatenate(pcms))


# ---------------------------------------------------------------------------------------------
# the C++ class API, driven like alacconvert drives libalac (one Encode() per frame)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["wav05_mono16_k0", "wav50_stereo16_k0", "music_stereo24_k0", "music_5ch16_k0"])
def test_class_api_demo(tmp_path, name):
    import alac_b200
    libdir = os.path.dirname(alac_b200.library_path())
    exe = str(tmp_path / "class_api_demo")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(_ROOT, "include"),
                    os.path.join(_ROOT, "tests", "cpp", "class_api_demo.cpp"), "-L" + libdir, "-lalac_b200",
                    "-Wl,-rpath," + libdir, "-o", exe], check=True)
    g = np.load(os.path.join(_ROOT, "tests", "golden", name + ".npz"))
    raw = str(tmp_path / "in.raw")
    g["pcm"].tofile(raw)
    prefix = str(tmp_path / "out")
    subprocess.run([exe, raw, str(int(g["channels"])), str(int(g["depth"])), str(int(g["sample_rate"])), prefix], check=True)
    assert open(prefix + ".cookie", "rb").read() == bytes(g["cookie"])
    assert np.array_equal(np.fromfile(prefix + ".sizes", np.uint32), g["sizes"])
    assert np.array_equal(np.fromfile(prefix + ".packets", np.uint8), g["packets"])
    assert np.array_equal(np.fromfile(prefix + ".pcm", np.uint8), g["pcm"])


# ---------------------------------------------------------------------------------------------
# alacconvert CLI: WAV -> CAF -> WAV through the batched ABI, files byte-compared (BASELINE config 1 shape)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["wav05_mono16_k0", "wav50_stereo16_k0", "music_stereo24_k0", "music_5ch16_k0"])
def test_alacconvert_cli(tmp_path, name):
    import alac_b200
    from tests import caf_ref
    exe = os.path.join(os.path.dirname(alac_b200.library_path()), "alacconvert")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.dirname(exe), "alacconvert"], check=True)
    g = np.load(os.path.join(_ROOT, "tests", "golden", name + ".npz"))
    ch, depth, sr = int(g["channels"]), int(g["depth"]), int(g["sample_rate"])
    wav, caf, back = str(tmp_path / "in.wav"), str(tmp_path / "out.caf"), str(tmp_path / "back.wav")
    open(wav, "wb").write(caf_ref.wav_bytes(sr, ch, depth, g["pcm"].tobytes()))
    subprocess.run([exe, wav, caf], check=True, stdout=subprocess.DEVNULL)
    want = caf_ref.caf_bytes(sr, ch, depth, bytes(g["cookie"]), int(g["pcm"].nbytes), g["packets"].tobytes(), g["sizes"])
    assert open(caf, "rb").read() == want
    subprocess.run([exe, caf, back], check=True, stdout=subprocess.DEVNULL)
    assert open(back, "rb").read() == open(wav, "rb").read()


def test_host_pipeline_matches_device_path(engine, oracle):
    """Host buffers large enough for several pipelined chunks on several compute streams: same bytes as
    the single-stream device-resident path, and decode is the identity (K = 1 and K = 3)."""
    import torch
    import alac_b200
    dev = torch.device("cuda", 0)
    for ch, depth, K in [(1, 16, 1), (2, 24, 3)]:
        frames = 4096 * 12000 + 777
        pcm_t = torch.cat([synth.corpus_torch(a, min(1 << 24, frames - a), ch, depth, dev, seed=3) for a in range(0, frames, 1 << 24)])
        cfg = alac_b200.EncoderConfig(channels=ch, bit_depth=depth, frames_per_segment=K)
        ref = engine.encode(pcm_t, cfg)
        pcm_h = pcm_t.cpu().numpy()
        got = engine.encode(pcm_h, cfg)
        assert got.num_packets == ref.num_packets > 8192
        assert np.array_equal(np.asarray(got.sizes, np.uint32), ref.sizes.cpu().numpy().astype(np.uint32))
        assert np.array_equal(got.packets, ref.packets.cpu().numpy())
        dec = engine.decode(got.cookie, got.packets, got.sizes)
        assert dec.status == 0 and np.array_equal(dec.pcm, pcm_h)
        _check = oracle.Encoder(ch, depth).encode_stream(pcm_h[: 4096 * 6 * cfg.bytes_per_frame], K)
        assert np.array_equal(got.packets[: _check.packets.nbytes], _check.packets)


def test_ber_table_on_device(engine):
    """Device-side parse of a CAF packet table (N3): sizes, stop at a zero entry, stop when the data runs out."""
    import torch
    from tests import caf_ref
    rng = np.random.default_rng(1)
    sizes = np.concatenate([rng.integers(1, 128, 300), rng.integers(128, 16384, 3000), rng.integers(16384, 200000, 500),
                            [127, 128, 16383, 16384, 2097151, 2097152]]).astype(np.uint32)
    rng.shuffle(sizes)
    table = np.frombuffer(b"".join(caf_ref.ber(int(s)) for s in sizes), np.uint8).copy()
    total = int(sizes.astype(np.int64).sum())
    got = engine.ber_table_sizes(table, total)
    assert np.array_equal(got, sizes)
    got_d = engine.ber_table_sizes(torch.from_numpy(table).cuda(), total)
    assert np.array_equal(got_d.cpu().numpy().astype(np.uint32), sizes)
    # trailing zero padding (the worst-case sized table of the writer) ends the list
    assert np.array_equal(engine.ber_table_sizes(np.concatenate([table, np.zeros(40, np.uint8)]), total), sizes)
    # data chunk shorter than the table claims: stop at the first packet that does not fit
    cut = int(sizes[:1000].astype(np.int64).sum()) + 5
    assert np.array_equal(engine.ber_table_sizes(table, cut), sizes[:1000])


def test_ber_table_build_on_device(engine):
    """sizes -> BER bytes on the device == the container writer's bytes; and back again."""
    import torch
    from tests import caf_ref
    rng = np.random.default_rng(2)
    sizes = np.concatenate([rng.integers(1, 128, 300), rng.integers(128, 16384, 3000), rng.integers(16384, 3000000, 500),
                            [1, 127, 128, 16383, 16384, 2097151, 2097152, 268435455, 268435456, 4294967295]]).astype(np.uint32)
    rng.shuffle(sizes)
    want = np.frombuffer(b"".join(caf_ref.ber(int(s)) for s in sizes), np.uint8)
    assert np.array_equal(engine.ber_table_build(sizes), want)
    got_d = engine.ber_table_build(torch.from_numpy(sizes.view(np.int32)).cuda())
    assert np.array_equal(got_d.cpu().numpy(), want)
    small = sizes[sizes < 200000]
    back = engine.ber_table_sizes(engine.ber_table_build(small), int(small.astype(np.int64).sum()))
    assert np.array_equal(back, small)


def test_decode_from_caf_table(engine):
    """Config-5 path: packets addressed through the BER table of a CAF file, all on the device."""
    import torch
    from tests import caf_ref
    g = np.load([p for p in _GOLDEN if "music_stereo32_k1" in p][0])
    table = np.frombuffer(b"".join(caf_ref.ber(int(s)) for s in g["sizes"]), np.uint8).copy()
    t_table, t_data = torch.from_numpy(table).cuda(), torch.from_numpy(g["packets"]).cuda()
    sizes = engine.ber_table_sizes(t_table, t_data.numel())
    dec = engine.decode(bytes(g["cookie"]), t_data, sizes)
    assert dec.status == 0 and np.array_equal(dec.pcm.cpu().numpy(), g["pcm"])


# ---------------------------------------------------------------------------------------------
# hand-crafted packets: syntax the encoder never emits but the decoder must accept
# (codec/ALACDecoder.cu:660-694 general predictor headers, :1012-1059 FIL / DSE)
# ---------------------------------------------------------------------------------------------
def _craft_sce(oracle, x, chan_bits, num, den_shift, mode, pb_factor, coefs, frame=4096, partial=None):
    """One SCE element holding x (int32, already within chan_bits): returns the element's bit list."""
    from tests.bitpack import BitWriter
    c = np.zeros(32, np.int16)
    c[:num] = coefs[:num]
    res = oracle.pc_block(x, c.copy(), num, chan_bits, den_shift)
    if mode:
        # the decoder undoes a first-difference pass (numactive 31) after dyn_decomp when mode != 0
        res = oracle.pc_block(res, np.zeros(32, np.int16), 31, chan_bits, 0)
    data, nbits, st = oracle.dyn_comp(res, chan_bits, pb=(40 * pb_factor) // 4)
    assert st == 0
    w = BitWriter()
    w.put(0, 3); w.put(0, 4); w.put(0, 12)                      # ID_SCE, instance tag, unused
    n = len(x)
    is_partial = n != frame
    w.put((int(is_partial) << 3) | 0, 4)                        # partial, bytesShifted = 0, escape = 0
    if is_partial:
        w.put(n, 32)
    w.put(0, 8); w.put(0, 8)                                    # mixBits, mixRes
    w.put((mode << 4) | den_shift, 8)
    w.put((pb_factor << 5) | num, 8)
    for k in range(num):
        w.put(int(c[k]) & 0xffff, 16)
    w.put_bytes(data, nbits)
    return w.bits


def _finish_packet(bit_list):
    from tests.bitpack import BitWriter
    w = BitWriter()
    w.bits = list(bit_list)
    w.put(7, 3)                                                 # ID_END
    return w.to_bytes()


@pytest.mark.parametrize("num,den_shift,mode,pb_factor", [(0, 9, 0, 4), (2, 9, 0, 4), (4, 8, 0, 4), (8, 9, 1, 4), (16, 9, 0, 4),
                                                          (31, 9, 0, 4), (12, 6, 1, 2), (8, 9, 0, 7), (5, 10, 0, 4)])
def test_decoder_general_predictor_headers(engine, oracle, num, den_shift, mode, pb_factor):
    """Orders other than 4 / 8, denShift != 9, mode != 0 and pbFactor != 4: GPU == oracle decoder == source."""
    import alac_b200
    rng = np.random.default_rng(num * 100 + den_shift)
    packets, sizes, want = [], [], []
    for i in range(40):                                         # more than one warp of packets, ragged lengths
        n = 4096 if i % 3 else int(rng.integers(40, 4096))
        x = np.cumsum(rng.integers(-300, 301, n)).astype(np.int64)
        x = (np.clip(x, -30000, 30000) + rng.integers(-40, 41, n)).astype(np.int32)
        coefs = (oracle.init_coefs(32, den_shift).astype(np.int32) >> (0 if num <= 8 else 1)).astype(np.int16)
        p = _finish_packet(_craft_sce(oracle, x, 16, num, den_shift, mode, pb_factor, coefs))
        packets.append(p); sizes.append(len(p)); want.append(x.astype(np.int16))
    cfg = alac_b200.EncoderConfig(channels=1, bit_depth=16)
    cookie = alac_b200.magic_cookie(cfg)
    blob, sz = np.concatenate(packets), np.array(sizes, np.uint32)
    ref, st = oracle.Decoder(cookie).decode_stream(blob, sz)
    assert not st.any()
    assert np.array_equal(ref.view(np.int16), np.concatenate(want)), "oracle decoder disagrees with the crafted source"
    dec = engine.decode(cookie, blob, sz)
    assert dec.status == 0 and np.array_equal(dec.pcm, ref)


def test_decoder_skips_fil_and_dse(engine, oracle):
    """FIL and DSE elements in front of / between audio elements are skipped (codec/ALACDecoder.cu:1012-1059)."""
    import alac_b200
    from tests.bitpack import BitWriter, bits_of
    g = np.load([p for p in _GOLDEN if "music_stereo16_k1" in p][0])
    sizes, packets = g["sizes"], g["packets"]
    rng = np.random.default_rng(5)
    out_p, out_s, off = [], [], 0
    for i, s in enumerate(sizes):
        body = bits_of(packets[off:off + int(s)])
        off += int(s)
        w = BitWriter()
        if i % 4 == 1:                                          # short FIL: count < 15
            cnt = int(rng.integers(0, 15))
            w.put(6, 3); w.put(cnt, 4)
            for _ in range(cnt): w.put(int(rng.integers(0, 256)), 8)
        elif i % 4 == 2:                                        # long FIL: count = 15 + ext - 1
            ext = int(rng.integers(1, 40))
            w.put(6, 3); w.put(15, 4); w.put(ext, 8)
            for _ in range(15 + ext - 1): w.put(int(rng.integers(0, 256)), 8)
        elif i % 4 == 3:                                        # DSE with byte alignment and a 255+ count
            cnt = 255 + int(rng.integers(0, 30))
            w.put(4, 3); w.put(3, 4); w.put(1, 1); w.put(255, 8); w.put(cnt - 255, 8)
            w.align()
            for _ in range(cnt): w.put(int(rng.integers(0, 256)), 8)
        w.bits.extend(body)                                     # the original elements + ID_END (+ padding)
        p = w.to_bytes()
        out_p.append(p); out_s.append(len(p))
    blob, sz = np.concatenate(out_p), np.array(out_s, np.uint32)
    ref, st = oracle.Decoder(bytes(g["cookie"])).decode_stream(blob, sz)
    assert not st.any() and np.array_equal(ref, g["pcm"])
    dec = engine.decode(bytes(g["cookie"]), blob, sz)
    assert dec.status == 0 and np.array_equal(dec.pcm, g["pcm"])


def test_odd_frame_length_and_mixed_kinds_in_one_group(engine, oracle):
    """A frame length that is not a multiple of the 32-sample tile, with escape (noise), run-heavy (silence) and
    ordinary packets decoded side by side in the same 32-packet groups; stereo 24-bit exercises the shift region."""
    import alac_b200
    ch, depth, F = 2, 24, 1000
    parts = []
    for i in range(70):
        kind = ["music", "noise", "silence", "music"][i % 4]
        parts.append(synth.make(kind, F, ch, depth, seed=100 + i))
    parts.append(synth.make("music", 333, ch, depth, seed=7))          # ragged tail frame
    pcm = np.concatenate(parts)
    got, _ = _check_encode(engine, oracle, pcm, ch, depth, K=1, frame_size=F)
    _check_decode(engine, oracle, got, pcm)


@pytest.mark.parametrize("seed", [11, 12, 13, 14, 15, 16])
def test_decode_fuzzed_packets_do_not_hang_or_leak(engine, oracle, seed):
    """Bit-flipped, truncated, header-damaged and random packets next to good ones: the call returns, the GPU and the
    oracle decoder agree on which packets are accepted (status 0 / kALAC_ParamErr), and every accepted packet decodes to
    the same samples.  Bytes past the end of a short packet read as zero on both sides."""
    import alac_b200
    rng = np.random.default_rng(seed)
    for ch, depth, golden in [(2, 16, "music_stereo16_k1"), (2, 24, "music_stereo24_k0"), (1, 20, "music_mono20_k2"),
                              (2, 32, "music_stereo32_k1"), (2, 16, "silence_stereo16_runs"), (2, 16, "noise_stereo16_escape")]:
        g = np.load([p for p in _GOLDEN if golden in p][0])
        sizes = g["sizes"].astype(np.int64)
        offs = np.concatenate([[0], np.cumsum(sizes)])
        good = [g["packets"][offs[i]:offs[i + 1]].copy() for i in range(len(sizes))]
        pk, is_good = [], []
        for rep in range(20):
            for i, p in enumerate(good):
                mode = int(rng.integers(0, 6))
                q = p.copy()
                if mode == 1:                                   # a few bit flips anywhere
                    for _ in range(int(rng.integers(1, 6))):
                        q[int(rng.integers(0, len(q)))] ^= 1 << int(rng.integers(0, 8))
                elif mode == 2:                                 # truncation
                    q = q[: int(rng.integers(1, len(q)))]
                elif mode == 3:                                 # header byte damage
                    q[int(rng.integers(0, min(8, len(q))))] = int(rng.integers(0, 256))
                elif mode == 4:                                 # noise
                    q = rng.integers(0, 256, int(rng.integers(1, 600)), dtype=np.uint8).astype(np.uint8)
                pk.append(q)
                is_good.append(mode in (0, 5))
        blob, sz = np.concatenate(pk), np.array([len(q) for q in pk], np.uint32)
        cookie = bytes(g["cookie"])
        dec = engine.decode(cookie, blob, sz, raise_on_error=False)
        ref = oracle.Decoder(cookie)
        F = 4096
        bpf = ch * (2 if depth == 16 else 4 if depth == 32 else 3)
        pos = 0
        st = np.asarray(dec.packet_status)
        ns = np.asarray(dec.packet_samples)
        checked = 0
        for i, q in enumerate(pk):
            n = int(ns[i])
            # (bytes past the end of a short packet read as zero in the oracle and in the CUDA bit readers alike)
            want, rst = ref.decode_stream(q, np.array([len(q)], np.uint32))
            if is_good[i]:
                assert st[i] == 0 and rst[0] == 0
            assert (st[i] == 0) == (rst[0] == 0), f"packet {i}: GPU status {st[i]}, oracle status {rst[0]}"
            if rst[0] == 0 and st[i] == 0:
                got = dec.pcm[pos * bpf:(pos + n) * bpf]
                assert len(want) == len(got) and np.array_equal(got, want), f"packet {i} differs"
                checked += 1
            pos += n
        assert checked >= sum(is_good)


@pytest.mark.parametrize("ch,depth", [(2, 16), (1, 16), (2, 24), (2, 32), (1, 20)])
def test_encode_pathological_pcm(engine, oracle, ch, depth):
    """PCM built to stress the integer corner cases: full-scale alternation (coefficient wrap, escape post-check),
    impulses in silence (zero runs of every length, run overflow past 65535 does not occur in a 4096 frame but
    long runs do), ramps and DC at the rails, random bytes; one signal per frame so that all of them share warps."""
    from tests.synth import pack
    rng = np.random.default_rng(depth * 10 + ch)
    F, hi, lo = 4096, (1 << (depth - 1)) - 1, -(1 << (depth - 1))
    t = np.arange(F)
    frames = [
        np.where(t % 2 == 0, hi, lo),                               # Nyquist at full scale
        np.where((t // 3) % 2 == 0, hi, lo),
        np.full(F, hi), np.full(F, lo), np.zeros(F, np.int64),
        np.where(t % 997 == 0, hi, 0),                              # sparse impulses
        np.where(t % 64 == 0, 1, 0), np.where(t % 5 == 0, -1, 0),
        np.clip((t - 2048) * (hi // 2048), lo, hi),                 # ramp rail to rail
        (t * 37) % 3 - 1,                                           # tiny dither around zero
        rng.integers(lo, hi + 1, F),                                # white noise
        np.cumsum(rng.integers(-3, 4, F)),                          # slow random walk
        np.where(t < 2048, 0, rng.integers(lo, hi + 1, F)),         # silence then noise
        np.where(t < 100, rng.integers(lo, hi + 1, F), 0),          # noise then silence
    ]
    sig = np.stack([np.concatenate(frames)] * ch, axis=1).astype(np.int64)
    if ch == 2:
        sig[:, 1] = np.roll(sig[:, 1], 7) // 2                      # a different but related second channel
    pcm = pack(sig, depth)
    for K in (1, 0):
        got, _ = _check_encode(engine, oracle, pcm, ch, depth, K=K)
        _check_decode(engine, oracle, got, pcm)
