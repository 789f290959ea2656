"""ipcl_python -- Python binding on the C ABI, shaped after the reference's
Python extension (intel/pailliercryptolib_python, referred to at
/root/reference/README.md:129-130).

The extension's sources are not part of /root/reference and there is no
network here, so this follows the class and method names of its published
usage: `PaillierKeypair.generate_keypair(n_length, enable_DJN)` returning a
`PaillierPublicKey` / `PaillierPrivateKey` pair, `encrypt` / `decrypt` on
scalars, lists and numpy arrays, `PaillierEncryptedNumber` with `+`, `*`,
`len()`, indexing and `sum()`, and `context.initializeContext`.  Integers
only (the extension's fixed-point float encoding is outside the modexp path).

Everything runs on the GPU through include/ipcl_b200.h: ciphertext batches are
device resident (`ipclb200_dev_*`), operators enqueue `*_dev` kernels on the
library stream, values come back to the host only in `decrypt()` (or
`ciphertexts()`).  Key generation follows ipcl/keygen.cpp:13-117 (DJN:
p = q = 3 mod 4, gcd(p-1, q-1) = 2, exact bit length, |p - q| > 2^(bits/2-100))
with the primality tests run as GPU modexp batches, as the C++ layer does
(ipcl/src/keygen.cpp).  There is no CPU fallback: without an sm_100 device the
first call raises IpclB200Error.
"""
import ctypes
import math
import secrets

import numpy as np

from . import capi
from .limbs import batch_from_limbs, batch_to_limbs, from_limbs, to_limbs

__all__ = ["PaillierKeypair", "PaillierPublicKey", "PaillierPrivateKey",
           "PaillierEncryptedNumber", "context"]


class context:  # noqa: N801  (name of the extension's module)
    """ipcl::initializeContext / terminateContext (ipcl/utils/context.cpp:40-86)"""

    @staticmethod
    def initializeContext(runtime_choice="GPU"):  # noqa: N802
        if runtime_choice.upper() not in ("DEFAULT", "CPU", "QAT", "HYBRID", "GPU", "B200"):
            raise ValueError("initializeContext: unknown runtime choice " + runtime_choice)
        capi.init(-1)
        return True

    @staticmethod
    def terminateContext():  # noqa: N802
        capi.shutdown()
        return True


class _DevBatch:
    """count x words little-endian limbs in HBM (ipclb200_dev_alloc)"""

    def __init__(self, count, words):
        self.count, self.words = int(count), int(words)
        p = ctypes.c_void_p()
        capi._check(capi.lib().ipclb200_dev_alloc(ctypes.c_size_t(self.nbytes), ctypes.byref(p)))
        self.ptr = p.value

    @property
    def nbytes(self):
        return self.count * self.words * 4

    @classmethod
    def from_numpy(cls, a):
        a = np.ascontiguousarray(a, dtype=np.uint32)
        b = cls(a.shape[0], a.shape[1])
        capi._check(capi.lib().ipclb200_dev_upload(
            ctypes.c_void_p(b.ptr), a.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(b.nbytes)))
        return b

    def to_numpy(self):
        out = np.empty((self.count, self.words), dtype=np.uint32)
        capi._check(capi.lib().ipclb200_dev_download(
            out.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(self.ptr),
            ctypes.c_size_t(self.nbytes)))
        return out

    def __del__(self):
        try:
            if self.ptr:
                capi.lib().ipclb200_dev_free(ctypes.c_void_p(self.ptr))
                self.ptr = None
        except Exception:  # interpreter shutdown
            pass


def _stream():
    capi.lib().ipclb200_stream.restype = ctypes.c_void_p
    return capi.lib().ipclb200_stream()


def _as_int_list(value):
    if isinstance(value, (int, np.integer)):
        return [int(value)], True
    if isinstance(value, np.ndarray):
        return [int(v) for v in value.reshape(-1)], False
    return [int(v) for v in value], False


class PaillierPublicKey:
    """ipcl::PublicKey (ipcl/pub_key.cpp): n, DJN constant hs, encrypt"""

    def __init__(self, n, n_length=None, enable_DJN=True, hs=None):  # noqa: N803
        self.n = int(n)
        self.n_length = int(n_length) if n_length else self.n.bit_length()
        self.nsquare = self.n * self.n
        self.nl = (self.n.bit_length() + 31) // 32
        if capi.lib().ipclb200_class_words(2 * self.nl) != 2 * self.nl:
            raise ValueError("key width %d bits is not a kernel size class" % self.n_length)
        self.enable_DJN = bool(enable_DJN)
        self.hs = None
        self.randbits = 0
        if self.enable_DJN:
            self.randbits = self.n_length >> 1
            self.hs = int(hs) if hs is not None else self._make_hs()
        self._key = capi.PubKey(to_limbs(self.n, self.nl),
                                None if self.hs is None else to_limbs(self.hs, 2 * self.nl),
                                self.randbits)

    def _make_hs(self):
        # hs = (-x^2)^n mod n^2, gcd(x, n) = 1 (ipcl/pub_key.cpp:32-49)
        while True:
            x = secrets.randbits(self.n.bit_length() + 128)
            if math.gcd(x, self.n) == 1:
                break
        h = (-(x % self.n) ** 2) % self.n
        W = 2 * self.nl
        out = capi.modexp(to_limbs(h, W)[None, :], to_limbs(self.n, self.nl)[None, :],
                          to_limbs(self.nsquare, W)[None, :], capi.SHARED_MOD)
        return from_limbs(out[0])

    def encrypt(self, value, apply_obfuscator=True):
        """PublicKey::encrypt (ipcl/pub_key.cpp:97-110); scalars, sequences or
        numpy arrays of non-negative integers (reduced mod n)"""
        vals, _ = _as_int_list(value)
        if not vals:
            raise ValueError("encrypt: Cannot encrypt empty PlainText")
        pt = _DevBatch.from_numpy(batch_to_limbs([v % self.n for v in vals], self.nl))
        ct = _DevBatch(len(vals), 2 * self.nl)
        r = None
        r_words = 0
        if apply_obfuscator:
            if self.enable_DJN:
                # the DJN randoms never exist on the host: one fresh 256-bit key and
                # 96-bit nonce from the OS, expanded in HBM by the ChaCha20 generator
                # (ipclb200_random_dev; getRandomBN per element in the reference,
                # ipcl/pub_key.cpp:59-61)
                r_words = (self.randbits + 31) // 32
                seed = np.frombuffer(secrets.token_bytes(44), dtype=np.uint32)
                r = _DevBatch(len(vals), r_words)
                capi.random_dev(r.ptr, len(vals), r_words, self.randbits, seed[:8], seed[8:],
                                0, _stream())
            else:
                r_words = self.nl
                rnd = batch_to_limbs([1 + secrets.randbelow(self.n - 1) for _ in vals], self.nl)
                r = _DevBatch.from_numpy(rnd)
        self._key.encrypt_dev(pt.ptr, self.nl, r.ptr if r else 0, r_words, len(vals),
                              ct.ptr, _stream(), make_secure=apply_obfuscator)
        return PaillierEncryptedNumber(self, ct)

    def __eq__(self, other):
        return isinstance(other, PaillierPublicKey) and self.n == other.n

    def __hash__(self):
        return hash(self.n)


class PaillierEncryptedNumber:
    """ipcl::CipherText (ipcl/ciphertext.cpp): a device-resident batch of
    ciphertexts bound to its public key, with the homomorphic operators"""

    def __init__(self, public_key, batch):
        self.public_key = public_key
        self._b = batch

    def __len__(self):
        return self._b.count

    def ciphertexts(self):
        """the ciphertexts as Python integers (downloads the batch)"""
        return batch_from_limbs(self._b.to_numpy())

    def __getitem__(self, idx):
        rows = self._b.to_numpy()[idx]
        return PaillierEncryptedNumber(self.public_key, _DevBatch.from_numpy(np.atleast_2d(rows)))

    def _mod_words(self):
        W = 2 * self.public_key.nl
        return to_limbs(self.public_key.nsquare, W), W

    def __add__(self, other):
        """ct + ct (a*b mod n^2, ipcl/ciphertext.cpp:35-69) or ct + plaintext
        (encode without obfuscator first, :75-80); a length-1 right operand is
        applied to every element"""
        if not isinstance(other, PaillierEncryptedNumber):
            other = self.public_key.encrypt(other, apply_obfuscator=False)
        if other.public_key != self.public_key:
            raise ValueError("CT + CT error: 2 different public keys detected!")
        if len(other) not in (len(self), 1):
            raise ValueError("CT + CT error: Size mismatch!")
        mod, W = self._mod_words()
        out = _DevBatch(len(self), W)
        flags = capi.SHARED_B if (len(other) == 1 and len(self) > 1) else 0
        capi.modmul_dev(self._b.ptr, other._b.ptr, mod, len(self), out.ptr, _stream(), flags)
        return PaillierEncryptedNumber(self.public_key, out)

    __radd__ = __add__

    def __mul__(self, other):
        """ct * plaintext = ct^pt mod n^2 (ipcl/ciphertext.cpp:83-106,143-162)"""
        if isinstance(other, PaillierEncryptedNumber):
            raise TypeError("CT * CT is not defined for Paillier")
        vals, scalar = _as_int_list(other)
        if len(vals) not in (len(self), 1):
            raise ValueError("CT * PT error: Size mismatch!")
        if any(v < 0 for v in vals):
            raise ValueError("ippModExp: negative exponent")
        ebits = max(1, max(v.bit_length() for v in vals))
        ew = (ebits + 31) // 32
        e = _DevBatch.from_numpy(batch_to_limbs(vals, ew))
        mod, W = self._mod_words()
        out = _DevBatch(len(self), W)
        flags = capi.SHARED_MOD | (capi.SHARED_EXP if len(vals) == 1 else 0)
        capi.modexp_dev(self._b.ptr, e.ptr, mod, ew, ebits, len(self), out.ptr, _stream(), flags)
        return PaillierEncryptedNumber(self.public_key, out)

    __rmul__ = __mul__

    def sum(self):
        """Enc(sum of all elements): a tree of ct + ct on halves"""
        cur = self
        while len(cur) > 1:
            rows = cur._b.to_numpy()
            if rows.shape[0] % 2:
                one = self.public_key.encrypt(0, apply_obfuscator=False)._b.to_numpy()
                rows = np.concatenate([rows, one])
            h = rows.shape[0] // 2
            a = PaillierEncryptedNumber(self.public_key, _DevBatch.from_numpy(rows[:h]))
            b = PaillierEncryptedNumber(self.public_key, _DevBatch.from_numpy(rows[h:]))
            cur = a + b
        return cur


class PaillierPrivateKey:
    """ipcl::PrivateKey (ipcl/pri_key.cpp): CRT decrypt by default"""

    def __init__(self, public_key, p, q):
        p, q = sorted((int(p), int(q)))
        if p * q != public_key.n:
            raise ValueError("PrivateKey ctor: Public key does not match p * q.")
        if p == q:
            raise ValueError("PrivateKey ctor: p and q are same")
        self.public_key = public_key
        self.p, self.q = p, q
        self.pl = (q.bit_length() + 31) // 32
        self._key = capi.PrivKey(to_limbs(p, self.pl), to_limbs(q, self.pl))
        self._crt = True

    def enableCRT(self, crt_on):  # noqa: N802
        self._crt = bool(crt_on)

    def decrypt(self, encrypted):
        """PrivateKey::decrypt (ipcl/pri_key.cpp:65-90); returns an int for a
        length-1 input, else a list of ints"""
        if encrypted.public_key != self.public_key:
            raise ValueError("decrypt: The value of N in public key mismatch.")
        if len(encrypted) == 0:
            raise ValueError("decrypt: Cannot decrypt empty CipherText")
        if encrypted._b.words != 4 * self.pl:
            raise ValueError("decrypt: ciphertext width does not match the key")
        out = _DevBatch(len(encrypted), 2 * self.pl)
        self._key.decrypt_dev(encrypted._b.ptr, len(encrypted), out.ptr, _stream(),
                              use_crt=self._crt)
        vals = batch_from_limbs(out.to_numpy())
        return vals[0] if len(vals) == 1 else vals


_SMALL_PRIMES = [p for p in range(3, 4096, 2)
                 if all(p % d for d in range(3, int(p ** 0.5) + 1, 2))]


def _prime_batch(bits, block=192, rounds=10):
    """one prime of exactly `bits` bits; Fermat base 2 on a sieved block and the
    Miller-Rabin rounds of the first survivor run as GPU modexp batches"""
    words = (bits + 31) // 32
    while True:
        cand = []
        while len(cand) < block:
            c = secrets.randbits(bits) | 1 | (1 << (bits - 1))
            if all(c % p for p in _SMALL_PRIMES):
                cand.append(c)
        mods = batch_to_limbs(cand, words)
        two = batch_to_limbs([2] * block, words)
        f = capi.modexp(two, batch_to_limbs([c - 1 for c in cand], words), mods)
        for i, c in enumerate(cand):
            if from_limbs(f[i]) != 1:
                continue
            d, s = c - 1, 0
            while d % 2 == 0:
                d //= 2
                s += 1
            bases = [2 + secrets.randbelow(c - 3) for _ in range(rounds)]
            x = capi.modexp(batch_to_limbs(bases, words), to_limbs(d, words)[None, :],
                            to_limbs(c, words)[None, :], capi.SHARED_EXP | capi.SHARED_MOD)
            ok = True
            for xi in batch_from_limbs(x):
                if xi in (1, c - 1):
                    continue
                for _ in range(s - 1):
                    xi = xi * xi % c
                    if xi == c - 1:
                        break
                else:
                    ok = False
                    break
            if ok:
                return c


class PaillierKeypair:
    """ipcl::generateKeypair (ipcl/keygen.cpp:92-117)"""

    @staticmethod
    def generate_keypair(n_length=2048, enable_DJN=True):  # noqa: N803
        if n_length > 4096:
            raise ValueError("generateKeyPair: modulus size in bits should belong to either "
                             "1Kb, 2Kb, 3Kb or 4Kb range only, key size exceed the range!!!")
        if n_length < 200 or n_length % 4:
            raise ValueError("generateKeyPair: key size should >=200, and divisible by 4")
        capi.init(-1)
        half = n_length // 2
        min_dist = 1 << (half - 100)
        while True:
            p = _prime_batch(half)
            q = _prime_batch(half)
            if p == q:
                continue
            if enable_DJN and (p % 4 != 3 or q % 4 != 3 or math.gcd(p - 1, q - 1) != 2):
                continue
            n = p * q
            if n.bit_length() != n_length or abs(p - q) <= min_dist:
                continue
            break
        pk = PaillierPublicKey(n, n_length, enable_DJN)
        return pk, PaillierPrivateKey(pk, p, q)
