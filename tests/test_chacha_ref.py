"""The oracle's ChaCha20 (oracle/chacha20_ref.py) against the test vector of
RFC 8439 section 2.3.2 and against the `cryptography` package."""
import os
import struct
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                "oracle"))
import chacha20_ref as cc  # noqa: E402

RFC_KEY = list(struct.unpack("<8I", bytes(range(32))))
RFC_NONCE = list(struct.unpack("<3I", bytes.fromhex("000000090000004a00000000")))
RFC_OUT = [0xe4e7f110, 0x15593bd1, 0x1fdd0f50, 0xc47120a3, 0xc7f4d1c7, 0x0368c033,
           0x9aaa2204, 0x4e6cd4c3, 0x466482d2, 0x09aa9f07, 0x05d7c214, 0xa2028bd9,
           0xd19c12b5, 0xb94e16de, 0xe883d0cb, 0x4e3c50a2]


def test_rfc8439_block_vector():
    assert cc.block(RFC_KEY, 1, RFC_NONCE) == RFC_OUT


def test_against_cryptography_package():
    algorithms = pytest.importorskip("cryptography.hazmat.primitives.ciphers.algorithms")
    from cryptography.hazmat.primitives.ciphers import Cipher
    rng = np.random.default_rng(11)
    key = rng.integers(0, 2**32, 8, dtype=np.uint64).astype(np.uint32)
    nonce = rng.integers(0, 2**32, 3, dtype=np.uint64).astype(np.uint32)
    first = 12345
    # `cryptography` takes counter (4 bytes LE) || nonce (12 bytes)
    full = struct.pack("<I", first) + nonce.tobytes()
    enc = Cipher(algorithms.ChaCha20(key.tobytes(), full), mode=None).encryptor()
    stream = np.frombuffer(enc.update(bytes(64 * 5)), dtype=np.uint32)
    mine = np.array([w for j in range(5) for w in cc.block(key, first + j, nonce)],
                    dtype=np.uint32)
    assert np.array_equal(stream, mine)


def test_batch_rule_truncates_to_bits():
    key, nonce = RFC_KEY, RFC_NONCE
    r = cc.batch_randoms(key, nonce, 3, 32, 1000, first_element=0)
    assert r.shape == (3, 32)
    assert (r[:, 31] >> 8 == 0).all() and r[:, 31].any()
    # element 0 of a 16-word batch starting at element 1 is block 1: the RFC vector
    r1 = cc.batch_randoms(key, nonce, 1, 16, 512, first_element=1)
    assert list(r1[0]) == RFC_OUT
