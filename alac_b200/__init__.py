"""alac_b200 -- B200-native ALAC codec engine (hand-written sm_100a kernels behind a C ABI).

The product is ``csrc/libalac_b200.so`` (see ``include/alac_b200.h``); this package is the thin
Python host layer over it: ctypes bindings, buffer plumbing for numpy (host) and torch CUDA
tensors (device), CAF container I/O and frame-range sharding across ranks.

There is no CPU codec path: importing works anywhere, but every encode/decode call needs the
built library and a CUDA device and raises otherwise.
"""
from .engine import (  # noqa: F401
    AlacError,
    EncodeResult,
    DecodeResult,
    Engine,
    EncoderConfig,
    library_path,
    load_library,
    magic_cookie,
    parse_cookie,
    encode_bound,
    Placement,
    DevicePtr,
    EXCHANGE_BYTES,
)

__all__ = [
    "AlacError", "EncodeResult", "DecodeResult", "Engine", "EncoderConfig",
    "library_path", "load_library", "magic_cookie", "parse_cookie", "encode_bound", "Placement", "DevicePtr", "EXCHANGE_BYTES",
]
