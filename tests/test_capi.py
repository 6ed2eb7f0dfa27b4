"""C-ABI checks that need no GPU: the library loads, exports every symbol include/alac_b200.h
declares, the host-only entry points work, and compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import alac_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for hdr in ("alac_b200.h", "alac_b200_container.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(alac_b200_[a-z_0-9]+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = alac_b200.load_library()
    names = _declared()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/alac_b200.h but not exported"


def test_class_api_symbols_exported():
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", alac_b200.library_path()], capture_output=True, text=True).stdout
    for sym in ["ALACEncoder17InitializeEncoder", "ALACEncoder6Encode", "ALACEncoder14GetMagicCookie",
                "ALACDecoder4Init", "ALACDecoder6Decode", "ALACEncoder11EncodeBatch", "ALACDecoder11DecodeBatch"]:
        assert sym in out, sym


def test_cookie_matches_oracle(oracle):
    for ch, depth, sr in [(1, 16, 44100), (2, 16, 44100), (2, 24, 96000), (8, 24, 48000), (6, 32, 48000)]:
        cfg = alac_b200.EncoderConfig(channels=ch, bit_depth=depth, sample_rate=sr)
        assert alac_b200.magic_cookie(cfg) == oracle.Encoder(ch, depth, sr).cookie()
        d = alac_b200.parse_cookie(alac_b200.magic_cookie(cfg))
        assert (d["num_channels"], d["bit_depth"], d["sample_rate"], d["frame_length"]) == (ch, depth, sr, 4096)
        assert (d["pb"], d["mb"], d["kb"], d["max_run"]) == (40, 10, 14, 255)


def test_cookie_wrappers_and_errors():
    cfg = alac_b200.EncoderConfig(channels=2, bit_depth=16)
    ck = alac_b200.magic_cookie(cfg)
    wrapped = bytes(4) + b"frma" + b"alac" + bytes(4) + b"alac" + bytes(4) + ck      # codec/ALACDecoder.cu:122-134
    assert alac_b200.parse_cookie(wrapped)["bit_depth"] == 16
    with pytest.raises(alac_b200.AlacError):
        alac_b200.parse_cookie(ck[:10])
    with pytest.raises(alac_b200.AlacError):
        alac_b200.magic_cookie(alac_b200.EncoderConfig(channels=9, bit_depth=16))


def test_encode_bound():
    cfg = alac_b200.EncoderConfig(channels=2, bit_depth=16)
    # one stereo packet: input bytes + kALACMaxEscapeHeaderBytes (codec/ALACAudioTypes.h:71) per packet
    assert alac_b200.encode_bound(cfg, 4096) == 4096 * 4 + 2 * 15
    assert alac_b200.encode_bound(cfg, 0) > 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(alac_b200.AlacError) as ei:
        alac_b200.Engine()
    assert ei.value.status == -1000


def test_product_does_not_import_oracle():
    """The product package must not reach into oracle/ (test infrastructure)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "alac_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f), errors="ignore").read().lower(), f
