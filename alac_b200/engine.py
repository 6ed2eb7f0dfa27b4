"""ctypes host layer over libalac_b200.so (the C ABI in include/alac_b200.h).

Mirrors the reference's encode/decode driver interface (convert-utility/main.cu:391-632 EncodeALAC,
:635-778 DecodeALAC) at batch granularity: one call takes a whole file or a batch of streams.
Buffers may be numpy arrays (host memory) or torch CUDA tensors (device memory); outputs live in
the same kind of memory as the inputs.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libalac_b200.so"

ALAC_OK = 0
ALAC_PARAM_ERROR = -50
ALAC_CUDA_ERROR = -1000
MEM_HOST, MEM_DEVICE = 0, 1
STATE_INT16S = 8 * 2 * 2 * 8
# torch reports its default stream as handle 0, which the C ABI reads as "the engine's own stream"; the explicit
# handle of CUDA's legacy default stream (cudaStreamLegacy) keeps the call ordered with torch's work
CUDA_STREAM_LEGACY = 1


class AlacError(RuntimeError):
    def __init__(self, status: int, what: str = ""):
        super().__init__(f"alac_b200 status {status}: {what}")
        self.status = status


def library_path() -> str:
    # ALAC_B200_LIB: developer override used to A/B-test kernel variants (still a CUDA build of the same ABI)
    return os.environ.get("ALAC_B200_LIB") or os.path.join(_HERE, "csrc", _LIB_NAME)


class _EncConfig(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("channels", C.c_uint32), ("bit_depth", C.c_uint32),
                ("frame_size", C.c_uint32), ("fast_mode", C.c_uint32), ("frames_per_segment", C.c_uint32)]


class _Stream(C.Structure):
    _fields_ = [("first_sample_frame", C.c_uint64), ("num_sample_frames", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [("num_packets", C.c_uint64), ("payload_bytes", C.c_uint64), ("escape_elements", C.c_uint64),
                ("max_packet_bytes", C.c_uint32), ("kernel_launches", C.c_uint32),
                ("ms_h2d", C.c_float), ("ms_kernels", C.c_float), ("ms_d2h", C.c_float),
                ("ms_search", C.c_float), ("ms_assemble", C.c_float), ("ms_decode", C.c_float),
                ("ms_final", C.c_float), ("ms_entropy", C.c_float), ("ms_finish", C.c_float), ("ms_fused", C.c_float),
                ("final_form", C.c_uint32), ("search_dense", C.c_uint32)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


class Placement(C.Structure):
    """alac_b200_placement (include/alac_b200.h): where one rank's packet block goes inside a job's single buffer."""
    _fields_ = [("dst_packets", C.c_void_p), ("dst_capacity", C.c_uint64), ("dst_sizes", C.c_void_p),
                ("first_packet", C.c_uint64), ("exchange", C.c_void_p), ("rank", C.c_uint32), ("n_ranks", C.c_uint32),
                ("home_rank", C.c_uint32), ("epoch", C.c_uint32), ("staging", C.c_void_p), ("slot_offsets", C.POINTER(C.c_uint64)), ("defer_finish", C.c_uint32)]


EXCHANGE_BYTES = 1024

_lib = None


def load_library():
    """Load libalac_b200.so.  Raises (never falls back) when the CUDA library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C alac_b200/csrc`. alac_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    vp, u32, i32, u64 = C.c_void_p, C.c_uint32, C.c_int32, C.c_uint64
    lib.alac_b200_version.restype = C.c_char_p
    lib.alac_b200_engine_create.argtypes = [i32, C.POINTER(vp)]
    lib.alac_b200_engine_create.restype = i32
    lib.alac_b200_engine_destroy.argtypes = [vp]
    lib.alac_b200_last_error.argtypes = [vp]
    lib.alac_b200_last_error.restype = C.c_char_p
    lib.alac_b200_engine_set_stream.argtypes = [vp, vp]
    lib.alac_b200_engine_set_stream.restype = i32
    lib.alac_b200_magic_cookie.argtypes = [C.POINTER(_EncConfig), u32, u32, vp, u32]
    lib.alac_b200_magic_cookie.restype = u32
    lib.alac_b200_encode_bound.argtypes = [C.POINTER(_EncConfig), u64, u64]
    lib.alac_b200_encode_bound.restype = u64
    lib.alac_b200_parse_cookie.argtypes = [vp, u32, C.POINTER(u32 * 11)]
    lib.alac_b200_parse_cookie.restype = i32
    lib.alac_b200_encode.argtypes = [vp, C.POINTER(_EncConfig), vp, u64, i32, C.POINTER(_Stream), u64,
                                     vp, u64, vp, u64, i32, vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(Stats)]
    lib.alac_b200_encode.restype = i32
    lib.alac_b200_decode.argtypes = [vp, vp, u32, vp, vp, u64, i32, vp, u64, vp, vp, i32, C.POINTER(u64), C.POINTER(Stats)]
    lib.alac_b200_decode.restype = i32
    lib.alac_b200_ber_table_sizes.argtypes = [vp, vp, u64, i32, u64, vp, u64, i32, C.POINTER(u64)]
    lib.alac_b200_ber_table_sizes.restype = i32
    lib.alac_b200_ber_table_build.argtypes = [vp, vp, u64, i32, vp, u64, i32, C.POINTER(u64)]
    lib.alac_b200_ber_table_build.restype = i32
    lib.alac_b200_encode_submit.argtypes = lib.alac_b200_encode.argtypes
    lib.alac_b200_encode_submit.restype = i32
    lib.alac_b200_decode_submit.argtypes = lib.alac_b200_decode.argtypes
    lib.alac_b200_decode_submit.restype = i32
    lib.alac_b200_wait.argtypes = [vp]
    lib.alac_b200_wait.restype = i32
    lib.alac_b200_engine_create_multi.argtypes = [C.POINTER(i32), u32, C.POINTER(vp)]
    lib.alac_b200_engine_create_multi.restype = i32
    lib.alac_b200_engine_num_devices.argtypes = [vp]
    lib.alac_b200_engine_num_devices.restype = u32
    lib.alac_b200_encode_placed.argtypes = [vp, C.POINTER(_EncConfig), vp, u64, i32, C.POINTER(_Stream), u64, C.POINTER(Placement),
                                            vp, u64, i32, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(vp), C.POINTER(Stats)]
    lib.alac_b200_encode_placed.restype = i32
    lib.alac_b200_placed_finish.argtypes = [vp, C.POINTER(u64)]
    lib.alac_b200_placed_finish.restype = i32
    lib.alac_b200_placed_base.argtypes = [vp, C.POINTER(u64)]
    lib.alac_b200_placed_base.restype = i32
    lib.alac_b200_device_alloc.argtypes = [vp, u64, C.POINTER(vp)]
    lib.alac_b200_device_alloc.restype = i32
    lib.alac_b200_device_free.argtypes = [vp, vp]
    lib.alac_b200_device_free.restype = i32
    lib.alac_b200_ipc_export.argtypes = [vp, vp, vp]
    lib.alac_b200_ipc_export.restype = i32
    lib.alac_b200_ipc_open.argtypes = [vp, vp, C.POINTER(vp)]
    lib.alac_b200_ipc_open.restype = i32
    lib.alac_b200_ipc_close.argtypes = [vp, vp]
    lib.alac_b200_ipc_close.restype = i32
    _lib = lib
    return lib


@dataclass
class EncoderConfig:
    """What InitializeEncoder/SetFrameSize/SetFastMode take (codec/ALACEncoder.cu:1457-1535) plus the
    encoder-reset schedule frames_per_segment (0 = one encoder for the whole stream)."""
    channels: int
    bit_depth: int
    sample_rate: int = 44100
    frame_size: int = 4096
    fast_mode: bool = False
    frames_per_segment: int = 1

    @property
    def bytes_per_frame(self) -> int:
        return {16: 2, 20: 3, 24: 3, 32: 4}[self.bit_depth] * self.channels

    def _c(self) -> _EncConfig:
        return _EncConfig(self.sample_rate, self.channels, self.bit_depth, self.frame_size,
                          int(self.fast_mode), self.frames_per_segment)


def magic_cookie(cfg: EncoderConfig, max_frame_bytes: int = 0, avg_bit_rate: int = 0) -> bytes:
    """ALACEncoder::GetMagicCookie (codec/ALACEncoder.cu:1109-1140)."""
    lib = load_library()
    buf = (C.c_uint8 * 48)()
    n = lib.alac_b200_magic_cookie(C.byref(cfg._c()), max_frame_bytes, avg_bit_rate, buf, 48)
    if n == 0:
        raise AlacError(ALAC_PARAM_ERROR, "bad encoder config")
    return bytes(buf[:n])


def parse_cookie(cookie: bytes) -> dict:
    lib = load_library()
    f = (C.c_uint32 * 11)()
    ck = (C.c_uint8 * len(cookie)).from_buffer_copy(cookie)
    st = lib.alac_b200_parse_cookie(ck, len(cookie), C.byref(f))
    if st:
        raise AlacError(st, "bad magic cookie")
    names = ["frame_length", "compatible_version", "bit_depth", "pb", "mb", "kb", "num_channels", "max_run",
             "max_frame_bytes", "avg_bit_rate", "sample_rate"]
    return dict(zip(names, list(f)))


def encode_bound(cfg: EncoderConfig, num_sample_frames: int, num_streams: int = 1) -> int:
    return int(load_library().alac_b200_encode_bound(C.byref(cfg._c()), num_sample_frames, num_streams))


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class DevicePtr:
    """A raw region of device memory (local, or another GPU's mapped by peer access / CUDA IPC): what a rank of a
    multi-GPU job passes when the bytes live in the job's shared buffer.  The caller orders the work itself."""

    def __init__(self, ptr: int, nbytes: int):
        self.ptr, self.nbytes = int(ptr), int(nbytes)

    def __getitem__(self, sl: slice) -> "DevicePtr":
        a, b, step = sl.indices(self.nbytes)
        if step != 1:
            raise ValueError("contiguous slices only")
        return DevicePtr(self.ptr + a, max(b - a, 0))


def _buf(x) -> Tuple[int, int, int]:
    """(pointer, nbytes, mem kind) of a numpy array, torch tensor or DevicePtr."""
    if isinstance(x, DevicePtr):
        return x.ptr, x.nbytes, MEM_DEVICE
    if _is_torch(x):
        if not x.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return x.data_ptr(), x.numel() * x.element_size(), (MEM_DEVICE if x.is_cuda else MEM_HOST)
    a = x
    if not isinstance(a, np.ndarray) or not a.flags["C_CONTIGUOUS"]:
        raise ValueError("expected a C-contiguous numpy array or torch tensor")
    return a.ctypes.data, a.nbytes, MEM_HOST


@dataclass
class EncodeResult:
    cookie: bytes
    packets: object           # uint8 numpy array / torch tensor, packets back to back (exactly `nbytes` long)
    sizes: object             # uint32 (numpy) / int32 (torch) per packet
    num_packets: int
    nbytes: int
    stats: dict = field(default_factory=dict)


@dataclass
class DecodeResult:
    pcm: object               # uint8, interleaved PCM
    sample_frames: int
    packet_samples: object
    packet_status: object
    status: int
    stats: dict = field(default_factory=dict)


class Engine:
    """One engine per GPU (owns a CUDA stream and its scratch)."""

    def __init__(self, device=-1):
        """device: a CUDA device index (-1 = current), or a list of indices for one engine that shards every call
        by frame range over several GPUs of this process (alac_b200_engine_create_multi; the first is the home device)."""
        self.lib = load_library()
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            arr = (C.c_int32 * len(device))(*device)
            st = self.lib.alac_b200_engine_create_multi(arr, len(device), C.byref(h))
        else:
            st = self.lib.alac_b200_engine_create(device, C.byref(h))
        if st:
            raise AlacError(st, "engine_create failed (no usable CUDA device?)")
        self.h = h

    @property
    def num_devices(self) -> int:
        return int(self.lib.alac_b200_engine_num_devices(self.h))

    # ------------------------------------------------------------------ shared device memory (one process per GPU)
    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        st = self.lib.alac_b200_device_alloc(self.h, nbytes, C.byref(p))
        if st:
            raise AlacError(st, self._err())
        return p.value

    def device_free(self, ptr: int):
        self.lib.alac_b200_device_free(self.h, C.c_void_p(ptr))

    def ipc_export(self, ptr: int) -> bytes:
        buf = (C.c_uint8 * 64)()
        st = self.lib.alac_b200_ipc_export(self.h, C.c_void_p(ptr), buf)
        if st:
            raise AlacError(st, self._err())
        return bytes(buf)

    def ipc_open(self, handle: bytes) -> int:
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        st = self.lib.alac_b200_ipc_open(self.h, buf, C.byref(p))
        if st:
            raise AlacError(st, self._err())
        return p.value

    def ipc_close(self, ptr: int):
        self.lib.alac_b200_ipc_close(self.h, C.c_void_p(ptr))

    def placed_finish(self) -> int:
        """Ends a staged job whose home-rank call used defer_finish (alac_b200_placed_finish); returns the job's bytes."""
        nb = C.c_uint64(0)
        st = self.lib.alac_b200_placed_finish(self.h, C.byref(nb))
        if st:
            raise AlacError(st, self._err())
        return nb.value

    def placed_base(self) -> int:
        """Byte offset of this rank's block inside the job's buffer (alac_b200_placed_base; after placed_finish)."""
        b = C.c_uint64(0)
        st = self.lib.alac_b200_placed_base(self.h, C.byref(b))
        if st:
            raise AlacError(st, "a deferred placed call is still pending")
        return b.value

    def encode_placed(self, pcm, cfg: EncoderConfig, placement: Placement, streams=None, out_sizes=None):
        """This rank's share of a job several GPUs encode together (alac_b200_encode_placed): the packets go straight
        to their final offset in the job's buffer (placement.dst_packets, usually another GPU's memory).  Returns
        (sizes of this rank's packets, num_packets, bytes, byte offset of the block, DevicePtr of the rank's own copy of
        the block, stats)."""
        ptr, nbytes, mem = _buf(pcm)
        self._follow_torch_stream(pcm)
        bpf = cfg.bytes_per_frame
        nsf = nbytes // bpf
        n_streams = len(streams) if streams is not None else 1
        ccfg = cfg._c()
        if streams is not None:
            max_packets = sum((n + cfg.frame_size - 1) // cfg.frame_size for _, n in streams)
            arr = (_Stream * n_streams)(*[_Stream(a, b) for a, b in streams])
        else:
            max_packets = (nsf + cfg.frame_size - 1) // cfg.frame_size
            arr = None
        max_packets = max(max_packets, 1)
        if out_sizes is None:
            if mem == MEM_DEVICE:
                import torch
                out_sizes = torch.empty(max_packets, dtype=torch.int32, device=pcm.device)
            else:
                out_sizes = np.empty(max_packets, np.uint32)
        sptr, scap, smem = _buf(out_sizes)
        npk, nb, base, stats, local = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), Stats(), C.c_void_p()
        st = self.lib.alac_b200_encode_placed(self.h, C.byref(ccfg), C.c_void_p(ptr), nsf, mem, arr, n_streams, C.byref(placement),
                                              C.c_void_p(sptr), scap // 4, smem, C.byref(npk), C.byref(nb), C.byref(base), C.byref(local),
                                              C.byref(stats))
        if st:
            raise AlacError(st, self._err())
        return out_sizes[:npk.value], npk.value, nb.value, base.value, DevicePtr(local.value or 0, nb.value), stats.as_dict()

    def close(self):
        if getattr(self, "h", None):
            self.lib.alac_b200_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self) -> str:
        return (self.lib.alac_b200_last_error(self.h) or b"").decode()

    def set_stream(self, cuda_stream_ptr: int = 0):
        self.lib.alac_b200_engine_set_stream(self.h, C.c_void_p(cuda_stream_ptr))

    def _follow_torch_stream(self, tensor):
        """Device-resident torch tensors are produced on torch's current stream: run the call on that stream
        so it is ordered after whatever filled them (and before whatever reads the results)."""
        if _is_torch(tensor) and tensor.is_cuda:
            import torch
            self.set_stream(torch.cuda.current_stream(tensor.device).cuda_stream or CUDA_STREAM_LEGACY)

    # ------------------------------------------------------------------ encode
    def encode_submit(self, pcm, cfg: EncoderConfig, **kw):
        """Asynchronous encode (alac_b200_encode_submit): returns a function that waits for the call and gives the
        EncodeResult.  One call per engine may be in flight; the buffers must stay alive until then."""
        return self.encode(pcm, cfg, _submit=True, **kw)

    def decode_submit(self, cookie: bytes, packets, sizes, **kw):
        """Asynchronous decode (alac_b200_decode_submit): returns a function that waits and gives the DecodeResult."""
        return self.decode(cookie, packets, sizes, _submit=True, **kw)

    def encode(self, pcm, cfg: EncoderConfig, streams: Optional[Sequence[Tuple[int, int]]] = None,
               coef_state: Optional[np.ndarray] = None, out=None, out_sizes=None, _submit: bool = False) -> EncodeResult:
        """Encode interleaved PCM (uint8 view; numpy = host, torch CUDA tensor = device).

        streams: optional [(first_sample_frame, num_sample_frames), ...]; default one stream over
        everything.  coef_state: optional int16 [n_streams, 256] carried encoder state (in/out).
        out / out_sizes: optional preallocated outputs (same memory kind as pcm).
        """
        ptr, nbytes, mem = _buf(pcm)
        self._follow_torch_stream(pcm)
        bpf = cfg.bytes_per_frame
        if nbytes % bpf:
            raise ValueError("pcm size is not a whole number of sample-frames")
        nsf = nbytes // bpf
        n_streams = len(streams) if streams is not None else 1
        ccfg = cfg._c()
        bound = int(self.lib.alac_b200_encode_bound(C.byref(ccfg), nsf, n_streams))
        if streams is not None:
            max_packets = sum((n + cfg.frame_size - 1) // cfg.frame_size for _, n in streams)
            arr = (_Stream * n_streams)(*[_Stream(a, b) for a, b in streams])
        else:
            max_packets = (nsf + cfg.frame_size - 1) // cfg.frame_size
            arr = None
        max_packets = max(max_packets, 1)
        if mem == MEM_DEVICE:
            import torch
            if out is None:
                out = torch.empty(bound, dtype=torch.uint8, device=pcm.device)
            if out_sizes is None:
                out_sizes = torch.empty(max_packets, dtype=torch.int32, device=pcm.device)
        else:
            if out is None:
                out = np.empty(bound, np.uint8)
            if out_sizes is None:
                out_sizes = np.empty(max_packets, np.uint32)
        optr, ocap, omem = _buf(out)
        sptr, scap, smem = _buf(out_sizes)
        if omem != mem or smem != mem:
            raise ValueError("outputs must live in the same memory kind as the input")
        state_ptr = None
        if coef_state is not None:
            if coef_state.dtype != np.int16 or coef_state.size != n_streams * STATE_INT16S or not coef_state.flags["C_CONTIGUOUS"]:
                raise ValueError("coef_state must be contiguous int16 [n_streams, 256]")
            state_ptr = coef_state.ctypes.data
        npk, nb, stats = C.c_uint64(0), C.c_uint64(0), Stats()
        fn = self.lib.alac_b200_encode_submit if _submit else self.lib.alac_b200_encode
        st = fn(self.h, C.byref(ccfg), C.c_void_p(ptr), nsf, mem, arr, n_streams,
                C.c_void_p(optr), ocap, C.c_void_p(sptr), scap // 4, mem,
                C.c_void_p(state_ptr) if state_ptr else None,
                C.byref(npk), C.byref(nb), C.byref(stats))
        if st:
            raise AlacError(st, self._err())

        def finish(wait: bool = True):
            if wait:
                rc = self.lib.alac_b200_wait(self.h)
                if rc:
                    raise AlacError(rc, self._err())
            keep = (pcm, arr, ccfg, coef_state)     # (alive until the call has ended)
            del keep
            return EncodeResult(magic_cookie(cfg), out[:nb.value], out_sizes[:npk.value], npk.value, nb.value, stats.as_dict())

        return finish if _submit else finish(False)

    # ------------------------------------------------------------------ CAF packet table
    def ber_table_sizes(self, table, data_bytes: int):
        """BER packet table of a CAF 'pakt' chunk (uint8 numpy array or torch CUDA tensor) -> packet sizes,
        parsed on the GPU.  Output lives where the table lives."""
        tptr, tbytes, mem = _buf(table)
        self._follow_torch_stream(table)
        if mem == MEM_DEVICE:
            import torch
            out = torch.empty(max(tbytes, 1), dtype=torch.int32, device=table.device)
        else:
            out = np.empty(max(tbytes, 1), np.uint32)
        n = C.c_uint64(0)
        st = self.lib.alac_b200_ber_table_sizes(self.h, C.c_void_p(tptr), tbytes, mem, data_bytes, C.c_void_p(_buf(out)[0]),
                                                max(tbytes, 1), mem, C.byref(n))
        if st:
            raise AlacError(st, self._err())
        return out[:n.value]

    def ber_table_build(self, sizes):
        """Packet sizes (uint32 numpy array or int32/uint32 torch CUDA tensor) -> BER bytes of a CAF 'pakt' chunk, built on
        the GPU.  Output lives where the sizes live."""
        sptr, sbytes, mem = _buf(sizes)
        self._follow_torch_stream(sizes)
        n = sbytes // 4
        if mem == MEM_DEVICE:
            import torch
            out = torch.empty(max(5 * n, 1), dtype=torch.uint8, device=sizes.device)
        else:
            out = np.empty(max(5 * n, 1), np.uint8)
        nb = C.c_uint64(0)
        st = self.lib.alac_b200_ber_table_build(self.h, C.c_void_p(sptr), n, mem, C.c_void_p(_buf(out)[0]), max(5 * n, 1), mem, C.byref(nb))
        if st:
            raise AlacError(st, self._err())
        return out[:nb.value]

    # ------------------------------------------------------------------ decode
    def decode(self, cookie: bytes, packets, sizes, out=None, raise_on_error: bool = True, _submit: bool = False) -> DecodeResult:
        """Decode packets laid back to back (uint8) with per-packet sizes (uint32/int32)."""
        cfgd = parse_cookie(cookie)
        bpf = {16: 2, 20: 3, 24: 3, 32: 4}[cfgd["bit_depth"]] * cfgd["num_channels"]
        pptr, pbytes, mem = _buf(packets)
        self._follow_torch_stream(sizes if isinstance(packets, DevicePtr) else packets)
        sptr, sbytes, smem = _buf(sizes)
        if smem != mem:
            raise ValueError("packets and sizes must live in the same memory kind")
        if (sizes.element_size() if _is_torch(sizes) else sizes.itemsize) != 4:
            raise ValueError("sizes must be 32-bit")
        n = sbytes // 4
        cap = n * cfgd["frame_length"] * bpf
        if mem == MEM_DEVICE:
            import torch
            where = sizes.device if isinstance(packets, DevicePtr) else packets.device
            if out is None:
                out = torch.empty(max(cap, 1), dtype=torch.uint8, device=where)
            psamp = torch.empty(max(n, 1), dtype=torch.int32, device=where)
            pstat = torch.empty(max(n, 1), dtype=torch.int32, device=where)
        else:
            if out is None:
                out = np.empty(max(cap, 1), np.uint8)
            psamp = np.zeros(max(n, 1), np.uint32)
            pstat = np.zeros(max(n, 1), np.int32)
        optr, ocap, omem = _buf(out)
        if omem != mem:
            raise ValueError("output must live in the same memory kind as the input")
        ck = (C.c_uint8 * len(cookie)).from_buffer_copy(cookie)
        nsf, stats = C.c_uint64(0), Stats()
        fn = self.lib.alac_b200_decode_submit if _submit else self.lib.alac_b200_decode
        st = fn(self.h, ck, len(cookie), C.c_void_p(pptr), C.c_void_p(sptr), n, mem,
                C.c_void_p(optr), ocap, C.c_void_p(_buf(psamp)[0]), C.c_void_p(_buf(pstat)[0]), mem,
                C.byref(nsf), C.byref(stats))
        if _submit and st:
            raise AlacError(st, self._err())

        def finish(rc):
            if rc not in (ALAC_OK, ALAC_PARAM_ERROR) or (rc and raise_on_error):
                raise AlacError(rc, self._err())
            keep = (packets, sizes, ck)             # (alive until the call has ended)
            del keep
            return DecodeResult(out[:nsf.value * bpf], nsf.value, psamp[:n], pstat[:n], rc, stats.as_dict())

        if _submit:
            return lambda: finish(self.lib.alac_b200_wait(self.h))
        return finish(st)
