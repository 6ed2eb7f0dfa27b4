"""MSB-first bit writer used by tests that hand-craft ALAC packets (codec/ALACBitUtilities.c BitBufferWrite semantics)."""
import numpy as np


class BitWriter:
    def __init__(self):
        self.bits = []

    def put(self, value: int, nbits: int):
        for i in range(nbits - 1, -1, -1):
            self.bits.append((value >> i) & 1)

    def put_bytes(self, data, nbits: int | None = None):
        """Append the first nbits (default: all) of a byte string, MSB first."""
        a = np.unpackbits(np.frombuffer(bytes(data), np.uint8))
        self.bits.extend(a[: len(a) if nbits is None else nbits].tolist())

    def align(self):
        while len(self.bits) % 8:
            self.bits.append(0)

    def to_bytes(self) -> np.ndarray:
        b = list(self.bits)
        while len(b) % 8:
            b.append(0)
        return np.packbits(np.array(b, np.uint8))


def bits_of(packet: np.ndarray) -> list:
    return np.unpackbits(np.asarray(packet, np.uint8)).tolist()
