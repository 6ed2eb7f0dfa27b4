"""CAF / WAV container I/O: the C library against an independent restatement of the
reference's file layout.  CPU only: packets come from the golden vectors."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import alac_b200
from tests import caf_ref

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


class PcmInfo(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("channels", C.c_uint32), ("bit_depth", C.c_uint32),
                ("data_offset", C.c_uint64), ("data_bytes", C.c_uint64)]


class CafInfo(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("channels", C.c_uint32), ("bit_depth", C.c_uint32), ("frames_per_packet", C.c_uint32),
                ("cookie", C.c_uint8 * 64), ("cookie_size", C.c_uint32), ("num_packets", C.c_uint64), ("valid_frames", C.c_uint64),
                ("table_offset", C.c_uint64), ("table_bytes", C.c_uint64), ("data_offset", C.c_uint64), ("data_bytes", C.c_uint64)]


@pytest.fixture(scope="module")
def lib():
    L = alac_b200.load_library()
    L.alac_b200_caf_write.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint64,
                                      C.c_void_p, C.c_void_p, C.c_uint64]
    L.alac_b200_caf_probe.argtypes = [C.c_char_p, C.POINTER(CafInfo)]
    L.alac_b200_caf_read_table.argtypes = [C.c_char_p, C.POINTER(CafInfo), C.c_void_p, C.c_uint64]
    L.alac_b200_caf_read_table.restype = C.c_uint64
    L.alac_b200_wav_probe.argtypes = [C.c_char_p, C.POINTER(PcmInfo)]
    L.alac_b200_wav_write.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64]
    L.alac_b200_ber_encode.argtypes = [C.c_uint32, C.c_void_p]
    L.alac_b200_ber_encode.restype = C.c_uint32
    L.alac_b200_ber_decode.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    L.alac_b200_ber_decode.restype = C.c_uint32
    return L


def test_ber_round_trip(lib):
    for v in [0, 1, 127, 128, 16383, 16384, 16392, 2097151, 2097152, 268435455, 268435456, 0xFFFFFFFF]:
        buf = (C.c_uint8 * 5)()
        n = lib.alac_b200_ber_encode(v, buf)
        assert bytes(buf[:n]) == caf_ref.ber(v)
        out = C.c_uint32(0)
        assert lib.alac_b200_ber_decode(buf, 5, C.byref(out)) == n and out.value == v


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_caf_write_matches_reference_layout(lib, tmp_path, path):
    g = np.load(path)
    ch, depth, sr = int(g["channels"]), int(g["depth"]), int(g["sample_rate"])
    cookie = bytes(g["cookie"])
    packets, sizes = np.ascontiguousarray(g["packets"]), np.ascontiguousarray(g["sizes"], np.uint32)
    out = str(tmp_path / "x.caf").encode()
    ck = (C.c_uint8 * len(cookie)).from_buffer_copy(cookie)
    st = lib.alac_b200_caf_write(out, sr, ch, depth, ck, len(cookie), int(g["pcm"].nbytes), packets.ctypes.data, sizes.ctypes.data, len(sizes))
    assert st == 0
    want = caf_ref.caf_bytes(sr, ch, depth, cookie, int(g["pcm"].nbytes), packets.tobytes(), sizes)
    assert open(out, "rb").read() == want
    info = CafInfo()
    assert lib.alac_b200_caf_probe(out, C.byref(info)) == 0
    assert (info.sample_rate, info.channels, info.bit_depth, info.frames_per_packet) == (sr, ch, depth, 4096)
    assert bytes(info.cookie[:info.cookie_size]) == cookie
    got = np.zeros(len(sizes) + 8, np.uint32)
    n = lib.alac_b200_caf_read_table(out, C.byref(info), got.ctypes.data, len(got))
    assert n == len(sizes) and np.array_equal(got[:n], sizes)
    raw = open(out, "rb").read()
    assert raw[info.data_offset:info.data_offset + packets.nbytes] == packets.tobytes()


def test_caf_exact_multiple_quirk(lib, tmp_path):
    """Length an exact multiple of 4096: the header counts one packet too many and says remainder 4096
    (CAFFileALAC.cpp:265-270) -- kept, because the files must be byte-identical."""
    cookie = alac_b200.magic_cookie(alac_b200.EncoderConfig(channels=2, bit_depth=16))
    sizes = np.array([100, 20000], np.uint32)
    packets = np.arange(20100, dtype=np.uint32).astype(np.uint8)
    out = str(tmp_path / "q.caf").encode()
    ck = (C.c_uint8 * len(cookie)).from_buffer_copy(cookie)
    assert lib.alac_b200_caf_write(out, 44100, 2, 16, ck, len(cookie), 2 * 4096 * 4, packets.ctypes.data, sizes.ctypes.data, 2) == 0
    raw = open(out, "rb").read()
    assert raw == caf_ref.caf_bytes(44100, 2, 16, cookie, 2 * 4096 * 4, packets.tobytes(), sizes)
    at = raw.index(b"pakt") + 12
    assert int.from_bytes(raw[at:at + 8], "big") == 3 and int.from_bytes(raw[at + 20:at + 24], "big") == 4096


def test_wav_write_and_probe(lib, tmp_path):
    for ch, depth in [(1, 16), (2, 16), (2, 24), (2, 32)]:
        pcm = np.arange(ch * (depth // 8 if depth != 20 else 3) * 1000, dtype=np.uint32).astype(np.uint8)
        out = str(tmp_path / f"w{ch}{depth}.wav").encode()
        assert lib.alac_b200_wav_write(out, 48000, ch, depth, pcm.ctypes.data, pcm.nbytes) == 0
        assert open(out, "rb").read() == caf_ref.wav_bytes(48000, ch, depth, pcm.tobytes())
        info = PcmInfo()
        assert lib.alac_b200_wav_probe(out, C.byref(info)) == 0
        assert (info.sample_rate, info.channels, info.bit_depth, info.data_offset, info.data_bytes) == (48000, ch, depth, 44, pcm.nbytes)


def test_probe_errors(lib, tmp_path):
    p = tmp_path / "junk.bin"
    p.write_bytes(b"not a container")
    assert lib.alac_b200_wav_probe(str(p).encode(), C.byref(PcmInfo())) == -50
    assert lib.alac_b200_caf_probe(str(p).encode(), C.byref(CafInfo())) == -50
    assert lib.alac_b200_wav_probe(str(tmp_path / "missing.wav").encode(), C.byref(PcmInfo())) == -43


def test_damaged_caf_is_rejected_not_walked_forever(lib, tmp_path):
    """Chunk sizes come from the file: a negative or huge size must end the walk with a parameter error (it used to
    seek backwards and loop), and a packet table that claims more bytes than the file holds reads nothing."""
    import struct
    d = np.load(GOLDEN[0])
    cookie = np.ascontiguousarray(d["cookie"]).astype(np.uint8)
    packets = np.ascontiguousarray(d["packets"]).astype(np.uint8)
    sizes = np.ascontiguousarray(d["sizes"]).astype(np.uint32)
    ch, depth = int(cookie[9]), int(cookie[5])
    good = str(tmp_path / "good.caf").encode()
    assert lib.alac_b200_caf_write(good, 44100, ch, depth, cookie.ctypes.data, cookie.size, 4096 * len(sizes) * ch * 2,
                                   packets.ctypes.data, sizes.ctypes.data, len(sizes)) == 0
    raw = bytearray(open(good, "rb").read())
    info = CafInfo()
    assert lib.alac_b200_caf_probe(good, C.byref(info)) == 0
    kuki = raw.index(b"kuki")
    for bad_size in (0xFFFFFFFFFFFFFFF0, 1 << 40, len(raw)):
        bad = bytearray(raw)
        bad[kuki + 4:kuki + 12] = struct.pack(">Q", bad_size)
        path = str(tmp_path / "bad.caf").encode()
        open(path, "wb").write(bad)
        assert lib.alac_b200_caf_probe(path, C.byref(CafInfo())) != 0
    lying = CafInfo.from_buffer_copy(bytes(info))
    lying.table_bytes = 1 << 40
    out = np.zeros(16, np.uint32)
    assert lib.alac_b200_caf_read_table(good, C.byref(lying), out.ctypes.data, out.size) == 0
    # a cookie whose frame length is not the 4096 that 'desc' and 'pakt' are written with
    c2 = cookie.copy()
    c2[0:4] = [0, 0, 8, 0]
    assert lib.alac_b200_caf_write(str(tmp_path / "x.caf").encode(), 44100, ch, depth, c2.ctypes.data, c2.size, 0, None, None, 0) != 0
