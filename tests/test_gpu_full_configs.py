"""BASELINE.json configs at their full sizes, through size-independent properties
(encode -> decode round trip is the identity; a sample of packets equals the CPU oracle;
sharded == unsharded).  Device-resident buffers throughout."""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu


def _corpus(first, frames, ch, depth, dev, seed=0):
    import torch
    step = 1 << 24
    return torch.cat([synth.corpus_torch(first + a, min(step, frames - a), ch, depth, dev, seed) for a in range(0, frames, step)])


def _oracle_check(oracle, cfg, pcm_t, enc, packets_to_check):
    """first / last `packets_to_check` packets byte-equal the oracle (K = 1: packets are independent)."""
    import torch
    bpf = cfg.bytes_per_frame
    F = cfg.frame_size
    sizes = enc.sizes.to(torch.int64)
    offs = torch.cat([torch.zeros(1, dtype=torch.int64, device=sizes.device), torch.cumsum(sizes, 0)])
    n = enc.num_packets
    for lo in (0, max(0, n - packets_to_check)):
        hi = min(n, lo + packets_to_check)
        pcm = pcm_t[lo * F * bpf: hi * F * bpf].cpu().numpy()
        want = oracle.Encoder(cfg.channels, cfg.bit_depth, cfg.sample_rate, reference=oracle.have_reference()).encode_stream(pcm, 1)
        got = enc.packets[int(offs[lo]): int(offs[hi])].cpu().numpy()
        assert np.array_equal(enc.sizes[lo:hi].cpu().numpy().astype(np.uint32), want.sizes)
        assert np.array_equal(got, want.packets)


def test_config2_one_hour_16_44_stereo(engine, oracle):
    import torch
    import alac_b200
    dev = torch.device("cuda", 0)
    cfg = alac_b200.EncoderConfig(channels=2, bit_depth=16, sample_rate=44100, frames_per_segment=1)
    pcm = _corpus(0, 3600 * 44100, 2, 16, dev)
    enc = engine.encode(pcm, cfg)
    assert enc.num_packets == 38760
    dec = engine.decode(enc.cookie, enc.packets, enc.sizes)
    assert dec.status == 0 and torch.equal(dec.pcm, pcm)
    _oracle_check(oracle, cfg, pcm, enc, 64)


def test_config3_ten_hours_24_96_stereo_sharded(engine, oracle):
    """10 h of 24-bit / 96 kHz stereo (20.7 GB): whole-corpus round trip on one GPU, and the frame-range shards
    of a 2/4/8-way split reproduce the same packets (checked on the shard that holds the tail)."""
    import torch
    import alac_b200
    from alac_b200 import shard
    dev = torch.device("cuda", 0)
    free, _ = torch.cuda.mem_get_info()
    hours = 10 if free > 150e9 else 2
    frames = hours * 3600 * 96000
    cfg = alac_b200.EncoderConfig(channels=2, bit_depth=24, sample_rate=96000, frames_per_segment=1)
    pcm = _corpus(0, frames, 2, 24, dev)
    enc = engine.encode(pcm, cfg)
    assert enc.num_packets == (frames + 4095) // 4096
    dec = engine.decode(enc.cookie, enc.packets, enc.sizes)
    assert dec.status == 0 and dec.sample_frames == frames and torch.equal(dec.pcm, pcm)
    del dec
    _oracle_check(oracle, cfg, pcm, enc, 16)
    sizes64 = enc.sizes.to(torch.int64)
    offs = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(sizes64, 0)])
    for world in (2, 4, 8):
        a, n = shard.plan_frame_shards(frames, 4096, world, 1)[world - 1]
        part = engine.encode(pcm[a * 6:(a + n) * 6], cfg)
        p0 = a // 4096
        assert torch.equal(part.sizes, enc.sizes[p0:])
        assert torch.equal(part.packets, enc.packets[int(offs[p0]):])


def test_config4_7_1_24_48(engine, oracle):
    import torch
    import alac_b200
    dev = torch.device("cuda", 0)
    cfg = alac_b200.EncoderConfig(channels=8, bit_depth=24, sample_rate=48000, frames_per_segment=1)
    pcm = _corpus(0, 600 * 48000, 8, 24, dev)
    enc = engine.encode(pcm, cfg)
    assert len(enc.cookie) == 48
    dec = engine.decode(enc.cookie, enc.packets, enc.sizes)
    assert dec.status == 0 and torch.equal(dec.pcm, pcm)
    _oracle_check(oracle, cfg, pcm, enc, 8)


def test_config5_100k_short_32bit_packets(engine, oracle):
    """100k packets of 32..4096 samples (32-bit stereo), each its own stream so every packet carries the
    partial-frame header; encoded in one multi-stream call, decoded in one call from the size table."""
    import torch
    import alac_b200
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(5)
    lens = rng.integers(32, 4097, size=100_000)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    total = int(lens.sum())
    cfg = alac_b200.EncoderConfig(channels=2, bit_depth=32, sample_rate=48000, frames_per_segment=0)
    pcm = _corpus(0, total, 2, 32, dev)
    enc = engine.encode(pcm, cfg, streams=[(int(a), int(n)) for a, n in zip(starts, lens)])
    assert enc.num_packets == 100_000
    dec = engine.decode(enc.cookie, enc.packets, enc.sizes)
    assert dec.status == 0 and dec.sample_frames == total
    assert torch.equal(dec.pcm, pcm)
    assert np.array_equal(dec.packet_samples.cpu().numpy().astype(np.int64), lens)
    # a few packets against the oracle
    sizes = enc.sizes.cpu().numpy().astype(np.int64)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    o = oracle.Encoder(2, 32, 48000, reference=oracle.have_reference())
    for i in [0, 1, 777, 99_999]:
        o.reset()
        a, n = int(starts[i]), int(lens[i])
        want = o.encode_packet(pcm[a * 8:(a + n) * 8].cpu().numpy(), n)
        assert np.array_equal(enc.packets[int(offs[i]):int(offs[i + 1])].cpu().numpy(), want)
