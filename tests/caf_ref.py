"""Independent pure-Python restatement of the reference's CAF / WAV file layout (checker for
alac_b200_container).  Follows SURVEY.md Appendix E: convert-utility/CAFFileALAC.cpp:60-286 and
convert-utility/main.cu:418-629, :803-852."""
import struct

LAYOUT_TAGS = [(100 << 16) | 1, (101 << 16) | 2, (113 << 16) | 3, (116 << 16) | 4,
               (120 << 16) | 5, (124 << 16) | 6, (142 << 16) | 7, (127 << 16) | 8]


def ber(v: int) -> bytes:
    groups = [v & 0x7F]
    v >>= 7
    while v:
        groups.append((v & 0x7F) | 0x80)
        v >>= 7
    return bytes(reversed(groups))


def caf_bytes(sample_rate, channels, depth, cookie: bytes, input_pcm_bytes: int, packets: bytes, sizes) -> bytes:
    bps = {16: 2, 20: 3, 24: 3, 32: 4}[depth]
    out = b"caff" + bytes([0, 1, 0, 0])
    flags = {16: 1, 20: 2, 24: 3, 32: 4}[depth]
    out += b"desc" + struct.pack(">q", 32) + struct.pack(">d", float(sample_rate)) + b"alac" + struct.pack(">IIIII", flags, 0, 4096, channels, 0)
    out += b"kuki" + struct.pack(">q", len(cookie)) + cookie
    if channels > 2:
        out += b"chan" + struct.pack(">q", 12) + struct.pack(">III", LAYOUT_TAGS[channels - 1], 0, 0)
    valid = input_pcm_bytes // (bps * channels)
    npk = valid // 4096
    rem = 4096 - (valid - npk * 4096)
    if rem:
        npk += 1
    entry = 2 if bps * channels * 4096 + 8 < 16384 else 3
    table_size = entry * npk
    table = b"".join(ber(int(s)) for s in sizes)
    left = table_size - len(table)
    hdr = struct.pack(">qqii", npk, valid, 0, rem)
    if left > 12:
        out += b"pakt" + struct.pack(">q", len(table) + 24) + hdr + table
        out += b"free" + struct.pack(">q", left - 12) + bytes(left - 12)
    else:
        out += b"pakt" + struct.pack(">q", table_size + 24) + hdr + table + bytes(left)
    out += b"data" + struct.pack(">q", len(packets) + 4) + struct.pack(">I", 1) + packets
    return out


def wav_bytes(sample_rate, channels, depth, pcm: bytes) -> bytes:
    block = channels * {16: 2, 20: 3, 24: 3, 32: 4}[depth]
    return (b"RIFF" + struct.pack("<I", len(pcm) + 36) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, channels, sample_rate,
            sample_rate * block, block, depth) + b"data" + struct.pack("<I", len(pcm)) + pcm)
