"""Driver-level parity against the reference EXECUTING: the fork's own unmodified `alacconvert`
(oracle/_ref/alacconvert_ref = /root/reference/codec/*.cu,*.c + convert-utility/main.cu,CAFFileALAC.cpp compiled
by oracle/Makefile) runs on the GPU box next to this repo's `alacconvert` (frames_per_segment = 0).

Asserted: the two .caf files are byte-identical (desc, kuki, pakt incl. the BER table, free, data) for the
reference's three WAV fixtures at FULL size (BASELINE config 1: audio/05.wav, 302 packets on one coefficient chain)
and for synthetic 16/24/32-bit mono and stereo inputs; and all four decode combinations (fork|ours decoder x
fork|ours file) give back the source PCM.  This pins EncodeStereo (codec/ALACEncoder.cu:290-558), EncodeMono
(:812-963), Encode (:973-1057), EncodeALAC (convert-utility/main.cu:391-632), Decode (codec/ALACDecoder.cu:571-1002)
and DecodeALAC (main.cu:635-778) as the reference runs them, not as anyone read them.

Inputs that hit defects of the fork are run too; what is asserted there is stated per case."""
import json
import os

import pytest

from tests import fork_pin

pytestmark = pytest.mark.gpu

_REPORT = {}


def _have_fork():
    return os.path.exists(fork_pin.FORK) and os.path.exists(fork_pin.KEEP)


def _pin(name, tmp_path):
    if not _have_fork():
        pytest.fail("oracle/_ref/alacconvert_ref is missing: run `make -C oracle` where /root/reference exists")
    if not os.path.exists(fork_pin.OURS):
        pytest.fail("alac_b200/csrc/alacconvert is missing: run __graft_entry__.build()")
    rep = fork_pin.pin_fixture(name, str(tmp_path))
    _REPORT[name] = rep
    out = os.path.join(fork_pin.ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "fork_pin_report.json"), "w") as f:
        json.dump(_REPORT, f, indent=1)
    return rep


@pytest.mark.parametrize("name", fork_pin.FIXTURES + list(fork_pin.SYNTH))
def test_caf_identical_to_the_forks_own_alacconvert(tmp_path, name):
    rep = _pin(name, tmp_path)
    assert rep["fork_encode_rc"] == 0 and rep["ours_encode_rc"] == 0
    enc = rep["encode"]
    assert enc["packets_ref"] == enc["packets_ours"]
    assert enc["packets_different"] == 0, enc["packet_diffs"][:3]
    for chunk in ("desc", "kuki", "pakt", "data"):
        assert enc[chunk + "_identical"], chunk
    assert enc["files_identical"]
    for leg, d in rep["decode"].items():
        assert d["rc"] == 0 and d["pcm_identical"], (leg, d)


def test_exact_multiple_of_the_frame_size(tmp_path):
    """The fork sizes its tables with X = bytes / packetBytes + 1 and reads outBytes[X-1] uninitialised when the
    input is an exact multiple of 4096 frames (convert-utility/main.cu:409,466 vs codec/ALACEncoder.cu:1389-1390).
    On this image the stray entry does not reach the file: the outputs still match byte for byte, including
    BuildBasePacketTable's extra packet count (CAFFileALAC.cpp:265-270)."""
    rep = _pin("exact_multiple_s16", tmp_path)
    assert rep["encode"]["files_identical"]
    assert all(d["pcm_identical"] for d in rep["decode"].values())


def test_escape_packets_where_the_fork_crashes(tmp_path):
    """Full-scale noise forces escape packets.  The fork's encoder dies there (SIGSEGV: EncodeStereoEscape reads a
    DEVICE pointer on the host, codec/ALACEncoder.cu:999 -> :770-775), so there is no fork file to compare; the fork's
    DECODER does run on this repo's escape packets and must give back the source PCM."""
    rep = _pin("noise_s16_escape", tmp_path)
    assert rep["ours_encode_rc"] == 0
    assert rep["fork_encode_rc"] != 0, "the fork no longer crashes on escape packets: compare the files instead"
    assert rep["decode"]["ours_decodes_ourfile"]["pcm_identical"]
    assert rep["decode"]["fork_decodes_ourfile"]["pcm_identical"], rep["decode"]["fork_decodes_ourfile"]
