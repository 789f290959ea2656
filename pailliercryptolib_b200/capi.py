"""ctypes binding of the C ABI in include/ipcl_b200.h.

This is the driver the parity tests and bench.py use to reach the library the
same way a foreign-language host would: plain pointers and sizes.  Arrays are
numpy uint32, little-endian limbs, shape (count, words).  There is no CPU
fallback here or below: if the CUDA library is missing or no sm_100 device is
present, calls raise."""
import ctypes
import os

import numpy as np

from . import build as _build

SHARED_BASE, SHARED_EXP, SHARED_MOD, SHARED_B = 1, 2, 4, 8

_u32p = ctypes.POINTER(ctypes.c_uint32)
_lib = None


class IpclB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ipcl_b200 error %d: %s" % (code, msg))
        self.code = code


EXPORTS = [
    "ipclb200_init", "ipclb200_shutdown", "ipclb200_device_count",
    "ipclb200_last_error", "ipclb200_version", "ipclb200_modexp",
    "ipclb200_modexp_dev", "ipclb200_modmul", "ipclb200_modmul_dev",
    "ipclb200_pubkey_create", "ipclb200_pubkey_destroy", "ipclb200_encrypt",
    "ipclb200_encrypt_dev", "ipclb200_privkey_create",
    "ipclb200_privkey_destroy", "ipclb200_decrypt", "ipclb200_decrypt_dev",
    "ipclb200_int_peak", "ipclb200_launch_count", "ipclb200_crt_residues",
    "ipclb200_pipe_mix", "ipclb200_stream", "ipclb200_dev_alloc",
    "ipclb200_dev_free", "ipclb200_dev_upload", "ipclb200_dev_download",
    "ipclb200_dev_copy", "ipclb200_sync", "ipclb200_class_words",
    "ipclb200_debug_montsqr", "ipclb200_int_peak_sustained",
    "ipclb200_init_devices", "ipclb200_active_devices", "ipclb200_has_experiments",
    "ipclb200_pubkey_set_table_policy", "ipclb200_batch_alloc", "ipclb200_batch_free",
    "ipclb200_batch_count", "ipclb200_batch_words", "ipclb200_batch_num_shards",
    "ipclb200_batch_shard", "ipclb200_batch_upload", "ipclb200_batch_download",
    "ipclb200_batch_sync", "ipclb200_batch_scatter", "ipclb200_batch_gather",
    "ipclb200_encrypt_batch", "ipclb200_decrypt_batch", "ipclb200_modmul_batch",
    "ipclb200_modexp_batch", "ipclb200_host_alloc", "ipclb200_host_free",
    "ipclb200_batch_touch", "ipclb200_privkey_set_schedule",
    "ipclb200_random_dev", "ipclb200_batch_random", "ipclb200_encrypt_drbg",
    "ipclb200_decrypt_layout", "ipclb200_zero_copy_count",
]


def lib():
    """Load libipcl_b200.so (must have been built in-tree; never built here
    on the fly on a GPU box without nvcc)."""
    global _lib
    if _lib is None:
        path = _build.CUDA_LIB
        if not os.path.exists(path):
            raise ImportError(
                "%s is missing: run `python -m pailliercryptolib_b200.build` "
                "(the CUDA extension is mandatory, there is no fallback)" % path)
        L = ctypes.CDLL(path)
        L.ipclb200_last_error.restype = ctypes.c_char_p
        L.ipclb200_version.restype = ctypes.c_char_p
        L.ipclb200_launch_count.restype = ctypes.c_uint64
        L.ipclb200_zero_copy_count.restype = ctypes.c_uint64
        L.ipclb200_batch_count.restype = ctypes.c_size_t
        L.ipclb200_stream.restype = ctypes.c_void_p
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise IpclB200Error(rc, lib().ipclb200_last_error().decode())


def _c(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _p(a):
    return None if a is None else a.ctypes.data_as(_u32p)


def _vp(x):
    return ctypes.c_void_p(int(x))


def init(device=-1):
    _check(lib().ipclb200_init(int(device)))


def init_devices(n=0):
    """one process, n GPUs (0 = all visible): batches are split over them"""
    _check(lib().ipclb200_init_devices(int(n)))
    return lib().ipclb200_active_devices()


def active_devices():
    return lib().ipclb200_active_devices()


def has_experiments():
    return bool(lib().ipclb200_has_experiments())


def shutdown():
    lib().ipclb200_shutdown()


def device_count():
    return lib().ipclb200_device_count()


def decrypt_layout(count, p_words, sms=0):
    """lane layout the CRT decrypt picks (host-only diagnostic)"""
    return int(lib().ipclb200_decrypt_layout(ctypes.c_size_t(count), int(p_words), int(sms)))


def launch_count():
    return int(lib().ipclb200_launch_count())


def zero_copy_count():
    return int(lib().ipclb200_zero_copy_count())


def int_peak():
    macs, mhz = ctypes.c_double(), ctypes.c_double()
    _check(lib().ipclb200_int_peak(ctypes.byref(macs), ctypes.byref(mhz)))
    return macs.value, mhz.value


def int_peak_sustained(seconds=0.3):
    macs = ctypes.c_double()
    _check(lib().ipclb200_int_peak_sustained(ctypes.c_double(seconds), ctypes.byref(macs)))
    return macs.value


def debug_montsqr(a, mod):
    a, mod = np.atleast_2d(_c(a)), _c(mod)
    assert a.shape[1] == 64 and mod.shape[-1] == 64
    s, m = np.zeros_like(a), np.zeros_like(a)
    _check(lib().ipclb200_debug_montsqr(_p(a), _p(mod), ctypes.c_size_t(a.shape[0]),
                                        _p(s), _p(m)))
    return s, m


def pipe_mix(mode):
    ms = ctypes.c_double()
    _check(lib().ipclb200_pipe_mix(int(mode), ctypes.byref(ms)))
    return ms.value


class _PinnedOwner:
    """keeps a page-locked allocation alive for the numpy array built on it"""

    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            if self.ptr and _lib is not None:
                _lib.ipclb200_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype=np.uint32):
    """numpy array in page-locked host memory (ipclb200_host_alloc).  The
    host-pointer entry points use such buffers in place (zero-copy over PCIe)
    for large batches; pageable arrays are staged by copies."""
    shape = tuple(int(x) for x in np.atleast_1d(shape))
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = ctypes.c_void_p()
    _check(lib().ipclb200_host_alloc(ctypes.c_size_t(max(nbytes, 16)), ctypes.byref(ptr)))
    owner = _PinnedOwner(ptr)
    buf = (ctypes.c_ubyte * max(nbytes, 16)).from_address(ptr.value)
    buf._owner = owner   # the ctypes buffer is the numpy array's base object
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def modexp(base, exp, mod, flags=0):
    base, exp, mod = _c(base), _c(exp), _c(mod)
    base2, exp2, mod2 = np.atleast_2d(base), np.atleast_2d(exp), np.atleast_2d(mod)
    count = max(base2.shape[0], exp2.shape[0], mod2.shape[0])
    mw, ew = mod2.shape[1], exp2.shape[1]
    assert base2.shape[1] == mw
    out = np.zeros((count, mw), dtype=np.uint32)
    _check(lib().ipclb200_modexp(_p(base2), _p(exp2), _p(mod2), mw, ew,
                                 ctypes.c_size_t(count), flags, _p(out)))
    return out


def modmul(a, b, mod, flags=0):
    a, b, mod = np.atleast_2d(_c(a)), np.atleast_2d(_c(b)), _c(mod)
    mw = mod.shape[-1]
    out = np.zeros_like(a)
    _check(lib().ipclb200_modmul(_p(a), _p(b), _p(mod), mw,
                                 ctypes.c_size_t(a.shape[0]), flags, _p(out)))
    return out


class PubKey:
    def __init__(self, n, hs=None, rand_bits=0):
        n = _c(n)
        self.n_words = n.shape[-1]
        self._h = ctypes.c_void_p()
        hs_a = None if hs is None else _c(hs)
        _check(lib().ipclb200_pubkey_create(_p(n), self.n_words, _p(hs_a),
                                            int(rand_bits), ctypes.byref(self._h)))

    def encrypt(self, pt, r=None, make_secure=True, out=None):
        pt = np.atleast_2d(_c(pt))
        r_a = None if r is None else np.atleast_2d(_c(r))
        ct = out if out is not None else np.zeros(
            (pt.shape[0], 2 * self.n_words), dtype=np.uint32)
        assert ct.dtype == np.uint32 and ct.flags.c_contiguous
        assert ct.shape == (pt.shape[0], 2 * self.n_words)
        _check(lib().ipclb200_encrypt(self._h, _p(pt), pt.shape[1], _p(r_a),
                                      0 if r_a is None else r_a.shape[1],
                                      ctypes.c_size_t(pt.shape[0]),
                                      int(make_secure), _p(ct)))
        return ct

    def encrypt_drbg(self, pt, key, nonce, out=None):
        """DJN encrypt with the r of the batch drawn on the device: ChaCha20
        keystream under key (8 words) / nonce (3 words)"""
        pt = np.atleast_2d(_c(pt))
        key, nonce = _c(key), _c(nonce)
        assert key.shape == (8,) and nonce.shape == (3,)
        ct = out if out is not None else np.zeros(
            (pt.shape[0], 2 * self.n_words), dtype=np.uint32)
        _check(lib().ipclb200_encrypt_drbg(self._h, _p(pt), pt.shape[1],
                                           ctypes.c_size_t(pt.shape[0]), _p(key),
                                           _p(nonce), _p(ct)))
        return ct

    def set_table_policy(self, max_table_mb=-1, upgrade_after=-1):
        _check(lib().ipclb200_pubkey_set_table_policy(
            self._h, ctypes.c_long(max_table_mb), ctypes.c_long(upgrade_after)))

    def encrypt_batch(self, pt, r, ct, r_bits=0, make_secure=True):
        _check(lib().ipclb200_encrypt_batch(self._h, pt._h, None if r is None else r._h,
                                            int(r_bits), int(make_secure), ct._h))

    def encrypt_dev(self, d_pt, pt_words, d_r, r_words, count, d_ct, stream,
                    make_secure=True):
        _check(lib().ipclb200_encrypt_dev(self._h, _vp(d_pt), pt_words, _vp(d_r),
                                          r_words, ctypes.c_size_t(count),
                                          int(make_secure), _vp(d_ct), _vp(stream)))

    def close(self):
        if self._h:
            lib().ipclb200_pubkey_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PrivKey:
    def __init__(self, p, q):
        p, q = _c(p), _c(q)
        self.p_words = p.shape[-1]
        self._h = ctypes.c_void_p()
        _check(lib().ipclb200_privkey_create(_p(p), _p(q), self.p_words,
                                             ctypes.byref(self._h)))

    def decrypt(self, ct, use_crt=True, out=None):
        ct = np.atleast_2d(_c(ct))
        assert ct.shape[1] == 4 * self.p_words
        pt = out if out is not None else np.zeros(
            (ct.shape[0], 2 * self.p_words), dtype=np.uint32)
        assert pt.dtype == np.uint32 and pt.flags.c_contiguous
        assert pt.shape == (ct.shape[0], 2 * self.p_words)
        _check(lib().ipclb200_decrypt(self._h, _p(ct), ctypes.c_size_t(ct.shape[0]),
                                      int(use_crt), _p(pt)))
        return pt

    def crt_residues(self, ct):
        """(count, 2, class words) array: ct^(p-1) mod p^2 and ct^(q-1) mod q^2
        as the decrypt kernel left them (diagnostics for the parity tests)"""
        ct = np.atleast_2d(_c(ct))
        assert ct.shape[1] == 4 * self.p_words
        xw = ctypes.c_int()
        _check(lib().ipclb200_crt_residues(self._h, _p(ct), ctypes.c_size_t(0),
                                           _p(ct), ctypes.byref(xw)))
        x = np.zeros((ct.shape[0], 2, xw.value), dtype=np.uint32)
        _check(lib().ipclb200_crt_residues(self._h, _p(ct),
                                           ctypes.c_size_t(ct.shape[0]), _p(x),
                                           ctypes.byref(xw)))
        return x

    def set_schedule(self, constant):
        """constant-schedule (fixed-window) ladders for the secret exponents"""
        _check(lib().ipclb200_privkey_set_schedule(self._h, int(bool(constant))))

    def decrypt_batch(self, ct, pt, use_crt=True):
        _check(lib().ipclb200_decrypt_batch(self._h, ct._h, int(use_crt), pt._h))

    def decrypt_dev(self, d_ct, count, d_pt, stream, use_crt=True):
        _check(lib().ipclb200_decrypt_dev(self._h, _vp(d_ct), ctypes.c_size_t(count),
                                          int(use_crt), _vp(d_pt), _vp(stream)))

    def close(self):
        if self._h:
            lib().ipclb200_privkey_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def modexp_dev(d_base, d_exp, mod, exp_words, exp_bits, count, d_out, stream,
               flags=SHARED_MOD):
    mod = _c(mod)
    _check(lib().ipclb200_modexp_dev(_vp(d_base), _vp(d_exp), _p(mod),
                                     mod.shape[-1], exp_words, exp_bits,
                                     ctypes.c_size_t(count), flags, _vp(d_out),
                                     _vp(stream)))


def random_dev(d_out, count, words, bits, key, nonce, first_element, stream):
    key, nonce = _c(key), _c(nonce)
    assert key.shape == (8,) and nonce.shape == (3,)
    _check(lib().ipclb200_random_dev(_vp(d_out), ctypes.c_size_t(count), int(words),
                                     int(bits), _p(key), _p(nonce),
                                     ctypes.c_uint64(first_element), _vp(stream)))


def modmul_dev(d_a, d_b, mod, count, d_out, stream, flags=0):
    mod = _c(mod)
    _check(lib().ipclb200_modmul_dev(_vp(d_a), _vp(d_b), _p(mod), mod.shape[-1],
                                     ctypes.c_size_t(count), flags, _vp(d_out),
                                     _vp(stream)))


class Batch:
    """count x words limbs in HBM, sharded over the active devices
    (ipclb200_batch_*)."""

    def __init__(self, count, words):
        self.count, self.words = int(count), int(words)
        self._h = ctypes.c_void_p()
        _check(lib().ipclb200_batch_alloc(ctypes.c_size_t(self.count), self.words,
                                          ctypes.byref(self._h)))

    @classmethod
    def from_host(cls, a, words=None):
        a = np.atleast_2d(_c(a))
        b = cls(a.shape[0], words or a.shape[1])
        b.upload(a)
        return b

    def upload(self, a):
        a = np.atleast_2d(_c(a))
        assert a.shape[0] == self.count
        _check(lib().ipclb200_batch_upload(self._h, _p(a), a.shape[1]))

    def download(self, words=None, out=None):
        w = words or self.words
        o = out if out is not None else np.zeros((self.count, w), dtype=np.uint32)
        _check(lib().ipclb200_batch_download(self._h, _p(o), w))
        return o

    def sync(self):
        _check(lib().ipclb200_batch_sync(self._h))

    def random(self, bits, key, nonce):
        """fill with ChaCha20 keystream truncated to `bits` bits per element"""
        key, nonce = _c(key), _c(nonce)
        assert key.shape == (8,) and nonce.shape == (3,)
        _check(lib().ipclb200_batch_random(self._h, int(bits), _p(key), _p(nonce)))

    def scatter_from(self, d_src):
        _check(lib().ipclb200_batch_scatter(self._h, _vp(d_src)))

    def gather_to(self, d_dst):
        _check(lib().ipclb200_batch_gather(self._h, _vp(d_dst)))

    def shards(self):
        out = []
        for i in range(lib().ipclb200_batch_num_shards(self._h)):
            dev, ptr = ctypes.c_int(), ctypes.c_void_p()
            begin, count = ctypes.c_size_t(), ctypes.c_size_t()
            stream = ctypes.c_void_p()
            _check(lib().ipclb200_batch_shard(self._h, i, ctypes.byref(dev), ctypes.byref(ptr),
                                              ctypes.byref(begin), ctypes.byref(count),
                                              ctypes.byref(stream)))
            out.append(dict(device=dev.value, ptr=ptr.value, begin=begin.value,
                            count=count.value, stream=stream.value))
        return out

    def close(self):
        if self._h:
            lib().ipclb200_batch_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def modmul_batch(a, b, mod, out, b_shared=None):
    mod = _c(mod)
    sh = None if b_shared is None else _c(b_shared)
    _check(lib().ipclb200_modmul_batch(a._h, None if b is None else b._h, _p(sh), _p(mod),
                                       mod.shape[-1], out._h))


def modexp_batch(base, exp, mod, out, exp_shared=None, exp_bits=0):
    mod = _c(mod)
    sh = None if exp_shared is None else _c(exp_shared)
    ew = exp.words if exp is not None else sh.shape[-1]
    _check(lib().ipclb200_modexp_batch(base._h, None if exp is None else exp._h, _p(sh), ew,
                                       int(exp_bits), _p(mod), mod.shape[-1], out._h))
