"""ctypes binding of the CPU oracle (oracle/liboracle.so, oracle/_ref/liboracle_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package alac_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "liboracle_ref.so")


def build(force: bool = False) -> None:
    """Compile the oracle (and, when /root/reference is present, oracle/_ref)."""
    if force or not os.path.exists(PORT_SO) or (os.path.isdir("/root/reference/codec") and not os.path.exists(REF_SO)):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)


class AgParams(C.Structure):
    _fields_ = [("mb0", C.c_uint32), ("pb", C.c_uint32), ("kb", C.c_uint32), ("wb", C.c_uint32)]


class Bits(C.Structure):
    _fields_ = [("buf", C.c_void_p), ("pos", C.c_uint64), ("cap", C.c_uint64)]


class Trace(C.Structure):
    _fields_ = [("tag", C.c_int32), ("escape", C.c_int32), ("mix_res", C.c_int32),
                ("num_u", C.c_int32), ("num_v", C.c_int32), ("bits_u", C.c_uint32), ("bits_v", C.c_uint32),
                ("hdr_coefs_u", C.c_int16 * 8), ("hdr_coefs_v", C.c_int16 * 8)]


class Config(C.Structure):
    _fields_ = [("frame_length", C.c_uint32), ("compatible_version", C.c_uint8), ("bit_depth", C.c_uint8),
                ("pb", C.c_uint8), ("mb", C.c_uint8), ("kb", C.c_uint8), ("num_channels", C.c_uint8),
                ("max_run", C.c_uint16), ("max_frame_bytes", C.c_uint32), ("avg_bit_rate", C.c_uint32),
                ("sample_rate", C.c_uint32)]


def _bind(lib):
    vp, u32, i32, u64 = C.c_void_p, C.c_uint32, C.c_int32, C.c_uint64
    lib.orc_init_coefs.argtypes = [vp, u32, i32]
    lib.orc_pc_block.argtypes = [vp, vp, i32, vp, i32, u32, u32]
    lib.orc_unpc_block.argtypes = [vp, vp, i32, vp, i32, u32, u32]
    lib.orc_ag_params_set.argtypes = [C.POINTER(AgParams), u32, u32, u32]
    lib.orc_dyn_comp.argtypes = [C.POINTER(AgParams), vp, C.POINTER(Bits), i32, i32, C.POINTER(u32)]
    lib.orc_dyn_comp.restype = i32
    lib.orc_dyn_decomp.argtypes = [C.POINTER(AgParams), C.POINTER(Bits), vp, i32, i32, C.POINTER(u32)]
    lib.orc_dyn_decomp.restype = i32
    lib.orc_prims_port.restype = vp
    lib.orc_prims_reference.restype = vp
    lib.orc_encoder_new.argtypes = [u32, u32, u32, u32, C.c_int, C.c_int]
    lib.orc_encoder_new.restype = vp
    lib.orc_encoder_free.argtypes = [vp]
    lib.orc_encoder_reset.argtypes = [vp]
    lib.orc_encoder_cookie.argtypes = [vp, vp, u32]
    lib.orc_encoder_cookie.restype = u32
    lib.orc_encode_packet.argtypes = [vp, vp, u32, vp, C.POINTER(u32), C.POINTER(Trace)]
    lib.orc_encode_packet.restype = i32
    lib.orc_encode_stream.argtypes = [vp, vp, u64, u32, vp, u64, C.POINTER(u64), vp, C.POINTER(u64)]
    lib.orc_encode_stream.restype = i32
    lib.orc_encoder_get_coefs.argtypes = [vp, u32, C.c_int, u32, vp]
    lib.orc_decoder_new.argtypes = [vp, u32, C.c_int, C.POINTER(i32)]
    lib.orc_decoder_new.restype = vp
    lib.orc_decoder_free.argtypes = [vp]
    lib.orc_decoder_config.argtypes = [vp]
    lib.orc_decoder_config.restype = C.POINTER(Config)
    lib.orc_decode_packet.argtypes = [vp, vp, u32, vp, C.POINTER(u32)]
    lib.orc_decode_packet.restype = i32
    lib.orc_decode_stream.argtypes = [vp, vp, vp, u64, vp, u64, C.POINTER(u64), vp]
    lib.orc_decode_stream.restype = i32
    lib.orc_fnv1a.argtypes = [vp, C.c_size_t]
    lib.orc_fnv1a.restype = u32
    return lib


_libs: dict = {}


def lib(reference: bool = False):
    """Load the oracle library; reference=True loads the flavour linked against the reference's primitives."""
    key = bool(reference)
    if key not in _libs:
        path = REF_SO if reference else PORT_SO
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        _libs[key] = _bind(C.CDLL(path))
    return _libs[key]


def have_reference() -> bool:
    try:
        return os.path.exists(REF_SO) or (os.path.isdir("/root/reference/codec") and (build() or os.path.exists(REF_SO)))
    except Exception:
        return False


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def bytes_per_sample(depth: int) -> int:
    return {16: 2, 20: 3, 24: 3, 32: 4}[depth]


# ---------------------------------------------------------------- primitives
def init_coefs(n: int = 16, denshift: int = 9) -> np.ndarray:
    c = np.zeros(n, np.int16)
    lib().orc_init_coefs(_ptr(c), denshift, n)
    return c


def _prim_lib(reference: bool):
    return lib(reference)


def pc_block(x: np.ndarray, coefs: np.ndarray, numactive: int, chanbits: int, denshift: int = 9,
             num: int | None = None, reference: bool = False) -> np.ndarray:
    """Returns residuals; coefs is updated in place.  reference=True runs the reference's pc_block."""
    x = np.ascontiguousarray(x, np.int32)
    num = len(x) if num is None else num
    pad = np.zeros(max(len(x), num) + 64, np.int32)
    pad[:len(x)] = x
    out = np.zeros_like(pad)
    L = _prim_lib(reference)
    if reference:
        fn = C.cast(C.c_void_p.from_address(L.orc_prims_reference() + 8).value,
                    C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_uint32, C.c_uint32))
        fn(_ptr(pad), _ptr(out), num, _ptr(coefs), numactive, chanbits, denshift)
    else:
        L.orc_pc_block(_ptr(pad), _ptr(out), num, _ptr(coefs), numactive, chanbits, denshift)
    return out[:max(num, 0)].copy()


def unpc_block(res: np.ndarray, coefs: np.ndarray, numactive: int, chanbits: int, denshift: int = 9,
               reference: bool = False) -> np.ndarray:
    res = np.ascontiguousarray(res, np.int32)
    num = len(res)
    pad = np.zeros(num + 64, np.int32)
    pad[:num] = res
    out = np.zeros_like(pad)
    L = _prim_lib(reference)
    if reference:
        fn = C.cast(C.c_void_p.from_address(L.orc_prims_reference() + 16).value,
                    C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_uint32, C.c_uint32))
        fn(_ptr(pad), _ptr(out), num, _ptr(coefs), numactive, chanbits, denshift)
    else:
        L.orc_unpc_block(_ptr(pad), _ptr(out), num, _ptr(coefs), numactive, chanbits, denshift)
    return out[:num].copy()


def dyn_comp(res: np.ndarray, bit_size: int, pb: int = 40, mb0: int = 10, kb: int = 14,
             reference: bool = False, start_bit: int = 0):
    """Returns (bytes, nbits, status)."""
    res = np.ascontiguousarray(res, np.int32)
    buf = np.zeros(len(res) * 5 + 64, np.uint8)
    L = _prim_lib(reference)
    ag = AgParams()
    L.orc_ag_params_set(C.byref(ag), mb0, pb, kb)
    b = Bits(buf.ctypes.data, start_bit, len(buf) * 8)
    nbits = C.c_uint32(0)
    if reference:
        fn = C.cast(C.c_void_p.from_address(L.orc_prims_reference() + 24).value,
                    C.CFUNCTYPE(C.c_int32, C.POINTER(AgParams), C.c_void_p, C.POINTER(Bits), C.c_int32, C.c_int32, C.POINTER(C.c_uint32)))
        st = fn(C.byref(ag), _ptr(res), C.byref(b), len(res), bit_size, C.byref(nbits))
    else:
        st = L.orc_dyn_comp(C.byref(ag), _ptr(res), C.byref(b), len(res), bit_size, C.byref(nbits))
    nbytes = (start_bit + nbits.value + 7) // 8
    return buf[:nbytes].copy(), nbits.value, st


def dyn_decomp(data: np.ndarray, num: int, max_size: int, pb: int = 40, mb0: int = 10, kb: int = 14,
               reference: bool = False, start_bit: int = 0, cap_bytes: int | None = None):
    """Returns (residuals, nbits, status)."""
    data = np.ascontiguousarray(data, np.uint8)
    buf = np.zeros(len(data) + 64, np.uint8)
    buf[:len(data)] = data
    out = np.zeros(num + 64, np.int32)
    L = _prim_lib(reference)
    ag = AgParams()
    L.orc_ag_params_set(C.byref(ag), mb0, pb, kb)
    b = Bits(buf.ctypes.data, start_bit, (len(data) if cap_bytes is None else cap_bytes) * 8)
    nbits = C.c_uint32(0)
    if reference:
        fn = C.cast(C.c_void_p.from_address(L.orc_prims_reference() + 32).value,
                    C.CFUNCTYPE(C.c_int32, C.POINTER(AgParams), C.POINTER(Bits), C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_uint32)))
        st = fn(C.byref(ag), C.byref(b), _ptr(out), num, max_size, C.byref(nbits))
    else:
        st = L.orc_dyn_decomp(C.byref(ag), C.byref(b), _ptr(out), num, max_size, C.byref(nbits))
    return out[:num].copy(), nbits.value, st


def fnv1a(a: np.ndarray) -> int:
    a = np.ascontiguousarray(a)
    return int(lib().orc_fnv1a(_ptr(a), a.nbytes))


# ---------------------------------------------------------------- codec objects
@dataclass
class EncodedStream:
    cookie: bytes
    packets: np.ndarray        # uint8, packets back to back
    sizes: np.ndarray          # uint32 per packet


class Encoder:
    """Oracle ALACEncoder (restated drivers; reference=True -> reference primitives)."""

    def __init__(self, channels: int, bit_depth: int, sample_rate: int = 44100, frame_size: int = 4096,
                 fast_mode: bool = False, reference: bool = False):
        self.L = lib(reference)
        self.channels, self.bit_depth, self.frame_size = channels, bit_depth, frame_size
        self.h = self.L.orc_encoder_new(channels, bit_depth, sample_rate, frame_size, int(fast_mode), int(reference))
        if not self.h:
            raise ValueError("orc_encoder_new failed")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_encoder_free(self.h)
            self.h = None

    @property
    def bytes_per_frame(self) -> int:
        return bytes_per_sample(self.bit_depth) * self.channels

    def reset(self):
        self.L.orc_encoder_reset(self.h)

    def cookie(self) -> bytes:
        buf = np.zeros(64, np.uint8)
        n = self.L.orc_encoder_cookie(self.h, _ptr(buf), 64)
        return bytes(buf[:n])

    def encode_packet(self, pcm: np.ndarray, num_samples: int, trace: bool = False):
        pcm = np.ascontiguousarray(pcm, np.uint8)
        assert pcm.nbytes >= num_samples * self.bytes_per_frame
        padded = np.zeros(pcm.nbytes + 64, np.uint8)   # tiny tails read a few entries past n*stride
        padded[:pcm.nbytes] = pcm
        out = np.zeros(self.frame_size * self.channels * 5 + 128, np.uint8)
        nb = C.c_uint32(0)
        tr = (Trace * 8)()
        st = self.L.orc_encode_packet(self.h, _ptr(padded), num_samples, _ptr(out), C.byref(nb), tr if trace else None)
        if st:
            raise RuntimeError(f"orc_encode_packet status {st}")
        return (out[:nb.value].copy(), list(tr)) if trace else out[:nb.value].copy()

    def encode_stream(self, pcm: np.ndarray, frames_per_segment: int = 0) -> EncodedStream:
        pcm = np.ascontiguousarray(pcm, np.uint8)
        nsf = pcm.nbytes // self.bytes_per_frame
        npk = (nsf + self.frame_size - 1) // self.frame_size
        cap = pcm.nbytes + 64 * (npk + 1)
        out = np.zeros(cap, np.uint8)
        sizes = np.zeros(max(npk, 1), np.uint32)
        ob, np_ = C.c_uint64(0), C.c_uint64(0)
        cookie = self.cookie()
        padded = np.zeros(pcm.nbytes + 64, np.uint8)
        padded[:pcm.nbytes] = pcm
        st = self.L.orc_encode_stream(self.h, _ptr(padded), nsf, frames_per_segment, _ptr(out), cap,
                                      C.byref(ob), _ptr(sizes), C.byref(np_))
        if st:
            raise RuntimeError(f"orc_encode_stream status {st}")
        return EncodedStream(cookie, out[:ob.value].copy(), sizes[:np_.value].copy())

    def coefs(self, channel: int, is_v: bool, row: int) -> np.ndarray:
        c = np.zeros(16, np.int16)
        self.L.orc_encoder_get_coefs(self.h, channel, int(is_v), row, _ptr(c))
        return c


class Decoder:
    def __init__(self, cookie: bytes, reference: bool = False):
        self.L = lib(reference)
        ck = np.frombuffer(cookie, np.uint8).copy()
        st = C.c_int32(0)
        self.h = self.L.orc_decoder_new(_ptr(ck), len(ck), int(reference), C.byref(st))
        if not self.h:
            raise ValueError(f"orc_decoder_new status {st.value}")
        self.cfg = self.L.orc_decoder_config(self.h).contents

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_decoder_free(self.h)
            self.h = None

    @property
    def bytes_per_frame(self) -> int:
        return bytes_per_sample(self.cfg.bit_depth) * self.cfg.num_channels

    def decode_packet(self, packet: np.ndarray):
        """Returns (pcm bytes, num_samples, status)."""
        packet = np.ascontiguousarray(packet, np.uint8)
        out = np.zeros(self.cfg.frame_length * self.bytes_per_frame + 64, np.uint8)
        n = C.c_uint32(0)
        st = self.L.orc_decode_packet(self.h, _ptr(packet), len(packet), _ptr(out), C.byref(n))
        return out[:n.value * self.bytes_per_frame].copy(), n.value, st

    def decode_stream(self, packets: np.ndarray, sizes: np.ndarray):
        """Returns (pcm bytes, statuses)."""
        packets = np.ascontiguousarray(packets, np.uint8)
        sizes = np.ascontiguousarray(sizes, np.uint32)
        cap = len(sizes) * self.cfg.frame_length * self.bytes_per_frame + 64
        out = np.zeros(cap, np.uint8)
        stat = np.zeros(max(len(sizes), 1), np.int32)
        nsf = C.c_uint64(0)
        self.L.orc_decode_stream(self.h, _ptr(packets), _ptr(sizes), len(sizes), _ptr(out), cap, C.byref(nsf), _ptr(stat))
        return out[:nsf.value * self.bytes_per_frame].copy(), stat[:len(sizes)].copy()
