"""Runs the FORK'S OWN, unmodified `alacconvert` (oracle/_ref/alacconvert_ref, built by oracle/Makefile from
/root/reference/codec/*.cu,*.c + convert-utility/main.cu,CAFFileALAC.cpp) next to this repo's `alacconvert`
on the reference's three WAV fixtures and compares the files chunk by chunk and packet by packet.

This is the reference *executing* its class drivers -- EncodeALAC (convert-utility/main.cu:391-632) ->
ALACEncoder::Encode (codec/ALACEncoder.cu:973-1057) -> EncodeStereo (:290-558) / EncodeMono (:812-963), and
DecodeALAC (main.cu:635-778) -> ALACDecoder::Decode (codec/ALACDecoder.cu:571-1002) -- so it pins the restated
drivers (oracle/alac_oracle.c and the CUDA kernels) to the reference itself, not to a second reading of it.

Test infrastructure only; needs a GPU (the fork segfaults without one)."""
import json
import os
import struct
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AUDIO = os.path.join(ROOT, "tests", "golden", "audio")
FORK = os.path.join(ROOT, "oracle", "_ref", "alacconvert_ref")
KEEP = os.path.join(ROOT, "oracle", "_ref", "libkeepfreed.so")
OURS = os.path.join(ROOT, "alac_b200", "csrc", "alacconvert")
FIXTURES = ["05.wav", "50.wav", "70.wav"]


def caf_chunks(blob: bytes):
    """[(type, body)] of a CAF file (CAFFileALAC.cpp:60-187 layout)."""
    assert blob[:4] == b"caff", "not a CAF file"
    out, pos = [], 8
    while pos + 12 <= len(blob):
        typ = blob[pos:pos + 4]
        size = struct.unpack(">q", blob[pos + 4:pos + 12])[0]
        body = blob[pos + 12:pos + 12 + size] if size >= 0 else blob[pos + 12:]
        out.append((typ.decode("latin1"), body))
        if size < 0:
            break
        pos += 12 + size
    return out


def ber_sizes(table: bytes, n: int):
    sizes, pos = [], 0
    while len(sizes) < n and pos < len(table):
        v = 0
        while True:
            b = table[pos]
            pos += 1
            v = (v << 7) | (b & 0x7F)
            if not b & 0x80:
                break
        sizes.append(v)
    return sizes, pos


def wav_data(blob: bytes) -> bytes:
    pos = 12
    while pos + 8 <= len(blob):
        tag, size = blob[pos:pos + 4], struct.unpack("<I", blob[pos + 4:pos + 8])[0]
        if tag == b"data":
            return blob[pos + 8:pos + 8 + size]
        pos += 8 + size + (size & 1)
    raise ValueError("no data chunk")


def run(cmd, env=None, cwd=None):
    try:
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env, cwd=cwd, timeout=120)
    except subprocess.TimeoutExpired:
        return -999, "timeout"
    return p.returncode, p.stdout.decode("latin1")


def compare_caf(ref: bytes, ours: bytes):
    """Chunk-by-chunk and packet-by-packet account of the differences between two ALAC CAF files."""
    rep = {"files_identical": ref == ours, "ref_bytes": len(ref), "ours_bytes": len(ours)}
    rc, oc = caf_chunks(ref), caf_chunks(ours)
    rep["chunk_order_ref"] = [t for t, _ in rc]
    rep["chunk_order_ours"] = [t for t, _ in oc]
    rd, od = dict(rc), dict(oc)
    for t in ("desc", "kuki", "pakt", "free", "data"):
        if t in rd or t in od:
            rep[t + "_identical"] = rd.get(t) == od.get(t)
    if "pakt" in rd and "pakt" in od and "data" in rd and "data" in od:
        rn = struct.unpack(">q", rd["pakt"][:8])[0]
        on = struct.unpack(">q", od["pakt"][:8])[0]
        rep["pakt_header_ref"] = struct.unpack(">qqii", rd["pakt"][:24])
        rep["pakt_header_ours"] = struct.unpack(">qqii", od["pakt"][:24])
        rs, _ = ber_sizes(rd["pakt"][24:], rn)
        os_, _ = ber_sizes(od["pakt"][24:], on)
        rep["packets_ref"], rep["packets_ours"] = len(rs), len(os_)
        diffs, rp, op = [], 4, 4
        for i in range(min(len(rs), len(os_))):
            a, b = rd["data"][rp:rp + rs[i]], od["data"][op:op + os_[i]]
            if a != b:
                first = next((j for j in range(min(len(a), len(b))) if a[j] != b[j]), min(len(a), len(b)))
                diffs.append({"packet": i, "ref_size": rs[i], "ours_size": os_[i], "first_diff_byte": first,
                              "ref_head": a[:24].hex(), "ours_head": b[:24].hex()})
            rp += rs[i]
            op += os_[i]
        rep["packets_different"] = len(diffs)
        rep["packet_diffs"] = diffs[:32]
    return rep


# Synthetic WAVs (made at test time from tests/synth.py) that take the fork through the depths and paths its own
# three 16-bit fixtures do not reach: the 24-/32-bit shift region (BASELINE config 3 is 24-bit / 96 kHz stereo),
# zero-run mode, and -- recorded rather than asserted -- the two inputs that hit defects of the fork.
SYNTH = {
    # name: (kind, sample-frames, channels, depth, rate)
    "music_s24_96k": ("music", 4096 * 30 + 1000, 2, 24, 96000),
    "corpus_s24_96k": ("corpus", 4096 * 20 + 77, 2, 24, 96000),
    "music_m24": ("music", 4096 * 12 + 3000, 1, 24, 48000),
    "music_s32": ("music", 4096 * 12 + 555, 2, 32, 44100),
    "music_m32": ("music", 4096 * 8 + 2048, 1, 32, 44100),
    "corpus_s16": ("corpus", 4096 * 40 + 1234, 2, 16, 44100),
    "silence_s16": ("silence", 4096 * 10 + 100, 2, 16, 44100),
    "silence_m16": ("silence", 4096 * 10 + 100, 1, 16, 44100),
}
# inputs on which the fork itself misbehaves; run, recorded in the report, never asserted identical
DEFECT = {
    "exact_multiple_s16": ("music", 4096 * 6, 2, 16, 44100),       # X off-by-one: outBytes[X-1] uninitialised (main.cu:409,466)
    "noise_s16_escape": ("noise", 4096 * 12 + 10, 2, 16, 44100),    # EncodeStereoEscape reads a device pointer on the host (ALACEncoder.cu:999,770)
}


def write_synth(name: str, workdir: str) -> str:
    import numpy as np
    from tests import synth
    from tests.caf_ref import wav_bytes
    kind, frames, ch, depth, rate = (SYNTH.get(name) or DEFECT[name])
    if kind == "corpus":
        import torch
        pcm = synth.corpus_torch(0, frames, ch, depth, torch.device("cpu")).numpy()
    else:
        pcm = synth.make(kind, frames, ch, depth, seed=11)
    path = os.path.join(workdir, name + ".wav")
    with open(path, "wb") as f:
        f.write(wav_bytes(rate, ch, depth, np.asarray(pcm, np.uint8).tobytes()))
    return path


def pin_fixture(name: str, workdir: str):
    """Encode `name` with both binaries, decode both ways, return the full report."""
    src = os.path.join(AUDIO, name) if name.endswith(".wav") else write_synth(name, workdir)
    stem = name.split(".")[0]
    ref_caf, our_caf = os.path.join(workdir, stem + "_ref.caf"), os.path.join(workdir, stem + "_ours.caf")
    rep = {"fixture": name}
    rep["fork_encode_rc"], rep["fork_encode_log"] = run([FORK, src, ref_caf])
    rep["ours_encode_rc"], _ = run([OURS, src, our_caf])
    if not (os.path.exists(ref_caf) and os.path.exists(our_caf)):
        rep["encode"] = {"files_identical": False, "missing_output": True}
        rep["decode"] = {}
        return rep
    ref, ours = open(ref_caf, "rb").read(), open(our_caf, "rb").read()
    try:
        rep["encode"] = compare_caf(ref, ours)
    except Exception as ex:      # a truncated / malformed fork output is a finding, not a crash of the checker
        rep["encode"] = {"files_identical": False, "malformed": repr(ex), "ref_bytes": len(ref), "ours_bytes": len(ours)}
    pcm = wav_data(open(src, "rb").read())
    # decode legs: 4 combinations (fork/ours decoder) x (fork/ours file)
    env = dict(os.environ, LD_PRELOAD=KEEP)
    dec = {}
    for dname, exe, e in (("fork", FORK, env), ("ours", OURS, None)):
        for fname, caf in (("forkfile", ref_caf), ("ourfile", our_caf)):
            out = os.path.join(workdir, f"{stem}_{dname}_{fname}.wav")
            rc, log = run([exe, caf, out], env=e)
            ok = False
            n = -1
            if os.path.exists(out):
                try:
                    got = wav_data(open(out, "rb").read())
                    n = len(got)
                    ok = got == pcm
                except Exception as ex:  # malformed output is a finding, not a crash
                    log += f"\n[parse] {ex}"
            dec[f"{dname}_decodes_{fname}"] = {"rc": rc, "pcm_identical": ok, "pcm_bytes": n, "want_bytes": len(pcm),
                                                "log_tail": log[-300:]}
    rep["decode"] = dec
    return rep


def main():
    import tempfile
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for name in FIXTURES + list(SYNTH) + list(DEFECT):
            out[name] = pin_fixture(name, d)
            print("done", name, out[name]["fork_encode_rc"], out[name]["encode"].get("files_identical"), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "fork_pin_report.json"), "w") as f:
        json.dump(out, f, indent=1)
    for name, r in out.items():
        e = r["encode"]
        print(name, "files_identical", e["files_identical"], "packets", e.get("packets_ref"), e.get("packets_ours"),
              "different", e.get("packets_different"), {k: v["pcm_identical"] for k, v in r["decode"].items()})


import sys
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

if __name__ == "__main__":
    main()
