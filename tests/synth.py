"""Deterministic synthetic PCM for parity tests and bench.py (integer recipes, numpy only)."""
from __future__ import annotations

import numpy as np

BYTES = {16: 2, 20: 3, 24: 3, 32: 4}


def pack(samples: np.ndarray, depth: int) -> np.ndarray:
    """int64/int32 [frames, channels] sample values (depth-bit signed) -> packed LE interleaved bytes."""
    s = np.ascontiguousarray(samples).astype(np.int64)
    lo, hi = -(1 << (depth - 1)), (1 << (depth - 1)) - 1
    s = np.clip(s, lo, hi)
    if depth == 16:
        return s.astype("<i2").view(np.uint8).reshape(-1).copy()
    if depth == 32:
        return s.astype("<i4").view(np.uint8).reshape(-1).copy()
    if depth == 20:
        s = s << 4                     # left-justified in 3 bytes
    u = (s & 0xFFFFFF).astype(np.uint32)
    out = np.empty(u.shape + (3,), np.uint8)
    out[..., 0] = u & 0xFF
    out[..., 1] = (u >> 8) & 0xFF
    out[..., 2] = (u >> 16) & 0xFF
    return out.reshape(-1).copy()


def unpack(raw: np.ndarray, depth: int, channels: int) -> np.ndarray:
    raw = np.ascontiguousarray(raw, np.uint8)
    if depth == 16:
        return raw.view("<i2").astype(np.int64).reshape(-1, channels)
    if depth == 32:
        return raw.view("<i4").astype(np.int64).reshape(-1, channels)
    b = raw.reshape(-1, 3).astype(np.int64)
    v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
    v = np.where(v & 0x800000, v - (1 << 24), v)
    if depth == 20:
        v = v >> 4
    return v.reshape(-1, channels)


def music(frames: int, channels: int, depth: int, seed: int = 1, rate: int = 44100) -> np.ndarray:
    """Tonal + noise floor, inter-channel correlated so different mixRes / numU,numV choices win."""
    rng = np.random.default_rng(seed)
    t = np.arange(frames, dtype=np.float64)
    full = float(1 << (depth - 1))
    base = np.zeros(frames)
    for k, (f, a) in enumerate([(220.0, 0.22), (554.37, 0.11), (1318.5, 0.05), (3520.0, 0.02)]):
        am = 1.0 + 0.3 * np.sin(2 * np.pi * (0.31 + 0.17 * k) * t / rate)
        base += a * am * np.sin(2 * np.pi * f * t / rate + k)
    out = np.empty((frames, channels), np.int64)
    noise_amp = max(2.0, full * 2.0 ** -9)
    for c in range(channels):
        side = 0.08 * np.sin(2 * np.pi * (97.0 + 41.0 * c) * t / rate + c) if c else 0.0
        gain = 1.0 - 0.11 * (c % 3)
        sig = full * (gain * base + side) + noise_amp * (rng.random(frames) - rng.random(frames))
        out[:, c] = np.round(sig).astype(np.int64)
    if depth == 32:
        out = (out & ~0xFFFF) | rng.integers(0, 1 << 16, size=out.shape)     # busy low 16 bits
    return out


def noise(frames: int, channels: int, depth: int, seed: int = 2) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.integers(-(1 << (depth - 1)), 1 << (depth - 1), size=(frames, channels), dtype=np.int64)


def silence_clicks(frames: int, channels: int, depth: int, seed: int = 3) -> np.ndarray:
    rng = np.random.default_rng(seed)
    out = np.zeros((frames, channels), np.int64)
    idx = rng.integers(0, frames, size=max(1, frames // 700))
    out[idx, :] = rng.integers(-200, 200, size=(len(idx), channels))
    return out


def square(frames: int, channels: int, depth: int, period: int = 37) -> np.ndarray:
    full = (1 << (depth - 1)) - 1
    t = np.arange(frames)
    s = np.where((t // period) % 2 == 0, full, -full - 1)
    return np.repeat(s[:, None], channels, axis=1).astype(np.int64)


KINDS = {"music": music, "noise": noise, "silence": silence_clicks, "square": square}


def make(kind: str, frames: int, channels: int, depth: int, seed: int = 1) -> np.ndarray:
    """Packed bytes of a synthetic signal."""
    fn = KINDS[kind]
    s = fn(frames, channels, depth, seed) if kind != "square" else fn(frames, channels, depth)
    return pack(s, depth)


# ---------------------------------------------------------------------------------------------
# Integer-only corpus generator (torch): identical bytes on CPU and on any GPU, so the CUDA arm,
# the CPU baseline and the reference arm of bench.py all see the same workload.
# ---------------------------------------------------------------------------------------------
def _hash32(t, salt: int):
    h = (t * 2654435761 + salt * 40503 + 0x9E3779B9) & 0xFFFFFFFF
    h = h ^ (h >> 15)
    h = (h * 2246822519) & 0xFFFFFFFF
    h = h ^ (h >> 13)
    h = (h * 3266489917) & 0xFFFFFFFF
    return h ^ (h >> 16)


def _para_sine(t, step: int, phase: int = 0):
    """Parabolic 'sine' of amplitude 2^15 from a 16-bit phase accumulator (integers only)."""
    x = ((t * step + phase) & 0xFFFF) - 32768
    return (x * (32768 - x.abs())) >> 13


def corpus_torch(first_frame: int, frames: int, channels: int, depth: int, device, seed: int = 0):
    """S-music corpus: per channel a sum of three detuned tones with slow AM plus a
    triangular-PDF noise floor; R = L - (L >> 3) + an independent low-level tone so the mid/side
    search is exercised.  Returns packed little-endian interleaved bytes as a torch.uint8 tensor."""
    import torch
    t = torch.arange(first_frame, first_frame + frames, dtype=torch.int64, device=device)
    am = 192 + (_para_sine(t >> 6, 3, seed * 977) >> 9)                     # 128..256, slow
    base = (_para_sine(t, 327 + 2 * seed) >> 2) + (_para_sine(t, 823, 1111) >> 3) + (_para_sine(t, 1961, 5000) >> 4)
    base = (base * am) >> 8                                                  # ~ +-13000 peak at 16 bit
    up = depth - 16 if depth != 32 else 0
    nz_bits = 7 if depth == 16 else 7 + min(up, 4)
    cols = []
    for c in range(channels):
        h = _hash32(t, 2 * c + 1 + 16 * seed)
        noise = (h & ((1 << nz_bits) - 1)) - ((h >> 12) & ((1 << nz_bits) - 1))
        if c % 2 == 0:
            s = base - ((base * (c // 2)) >> 3)
        else:
            s = base - (base >> 3) + (_para_sine(t, 97 + 41 * c, 300 * c) >> 5)
        s = (s << up) + noise if up else s + noise
        if depth == 32:
            s = (s << 16) | (_hash32(t, 99 + c) & 0xFFFF)                    # 16 busy low bits
        cols.append(s)
    x = torch.stack(cols, dim=1)                                             # [frames, channels] int64
    if depth == 16:
        return x.to(torch.int16).view(torch.uint8).reshape(-1)
    if depth == 32:
        return x.to(torch.int32).view(torch.uint8).reshape(-1)
    if depth == 20:
        x = x << 4
    b = torch.stack([(x & 0xFF), ((x >> 8) & 0xFF), ((x >> 16) & 0xFF)], dim=2).to(torch.uint8)
    return b.reshape(-1)
