"""Python int <-> little-endian uint32 limb arrays (the layout of the C ABI)."""
import numpy as np


def to_limbs(x, words):
    x = int(x)
    assert x >= 0 and x < (1 << (32 * words)), "value does not fit"
    return np.frombuffer(x.to_bytes(4 * words, "little"), dtype="<u4").astype(np.uint32)


def from_limbs(a):
    return int.from_bytes(np.ascontiguousarray(a, dtype="<u4").tobytes(), "little")


def batch_to_limbs(xs, words):
    out = np.zeros((len(xs), words), dtype=np.uint32)
    for i, x in enumerate(xs):
        out[i] = to_limbs(x, words)
    return out


def batch_from_limbs(a):
    a = np.ascontiguousarray(a, dtype="<u4")
    return [int.from_bytes(a[i].tobytes(), "little") for i in range(a.shape[0])]


def random_limbs(rng, count, words, top_mask=0xFFFFFFFF):
    """uniform random limb matrix from a numpy Generator"""
    a = rng.integers(0, 1 << 32, size=(count, words), dtype=np.uint64).astype(np.uint32)
    a[:, -1] &= np.uint32(top_mask)
    return a
