"""ORACLE (test infrastructure, never imported by the product): the ChaCha20
block function of RFC 8439 section 2.3 in plain Python, and the rule by which
the library turns its keystream into the DJN randoms of a batch
(chacha20_fill_kernel, pailliercryptolib_b200/csrc/kernels.cuh K6;
include/ipcl_b200.h "DJN randoms drawn on the device").

The reference itself draws r with getRandomBN(randbits) per element
(ipcl/pub_key.cpp:59-61, ipcl/utils/common.cpp:42-101); any uniform r is a
valid one, so what parity pins here is (a) the generator is exactly RFC 8439's
(test vector of section 2.3.2, plus the `cryptography` package as a second
opinion) and (b) encrypt with device-drawn r equals the oracle's encrypt with
the same r."""
import numpy as np

MASK = 0xFFFFFFFF


def _rotl(v, n):
    return ((v << n) & MASK) | (v >> (32 - n))


def _qr(x, a, b, c, d):
    x[a] = (x[a] + x[b]) & MASK; x[d] = _rotl(x[d] ^ x[a], 16)
    x[c] = (x[c] + x[d]) & MASK; x[b] = _rotl(x[b] ^ x[c], 12)
    x[a] = (x[a] + x[b]) & MASK; x[d] = _rotl(x[d] ^ x[a], 8)
    x[c] = (x[c] + x[d]) & MASK; x[b] = _rotl(x[b] ^ x[c], 7)


def block(key, counter, nonce):
    """key: 8 words, nonce: 3 words, counter: 32-bit -> the 16 output words"""
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + [int(k) for k in key] + \
         [counter & MASK] + [int(v) for v in nonce]
    x = list(st)
    for _ in range(10):
        _qr(x, 0, 4, 8, 12); _qr(x, 1, 5, 9, 13); _qr(x, 2, 6, 10, 14); _qr(x, 3, 7, 11, 15)
        _qr(x, 0, 5, 10, 15); _qr(x, 1, 6, 11, 12); _qr(x, 2, 7, 8, 13); _qr(x, 3, 4, 9, 14)
    return [(a + b) & MASK for a, b in zip(x, st)]


def batch_randoms(key, nonce, count, words, bits, first_element=0):
    """(count, words) uint32: element e = blocks e*bpe .. e*bpe+bpe-1 of the
    keystream, truncated to `bits` bits"""
    bpe = (words + 15) // 16
    out = np.zeros((count, words), dtype=np.uint32)
    for e in range(count):
        ws = []
        for j in range(bpe):
            ws += block(key, (first_element + e) * bpe + j, nonce)
        for w in range(words):
            left = bits - 32 * w
            v = ws[w]
            if left <= 0:
                v = 0
            elif left < 32:
                v &= (1 << left) - 1
            out[e, w] = v
    return out
