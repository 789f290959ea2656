"""ctypes front-end of the CPU oracle (oracle/paillier_oracle.c).

TEST INFRASTRUCTURE: import only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libpaillier_oracle.so")
_lib = None

_u32p = ctypes.POINTER(ctypes.c_uint32)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            import subprocess
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        _lib = ctypes.CDLL(_LIB)
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_u32p)


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint32)


def num_threads():
    return lib().orc_num_threads()


def modexp(base, exp, mod, shared_mod=False, shared_base=False,
           shared_exp=False):
    """base (count|1, L), exp (count|1, EL), mod (count|1, L) -> (count, L)."""
    base, exp, mod = _c(base), _c(exp), _c(mod)
    L, EL = mod.shape[-1], exp.shape[-1]
    count = max(base.shape[0], exp.shape[0], mod.shape[0])
    out = np.zeros((count, L), dtype=np.uint32)
    rc = lib().orc_modexp(_p(base), _p(exp), _p(mod), L, EL,
                          ctypes.c_size_t(count), int(shared_mod),
                          int(shared_base), int(shared_exp), _p(out))
    if rc:
        raise ValueError("orc_modexp rc=%d" % rc)
    return out


def modmul(a, b, mod, b_shared=False):
    a, b, mod = _c(a), _c(b), _c(mod)
    L = mod.shape[-1]
    out = np.zeros_like(a)
    rc = lib().orc_modmul(_p(a), _p(b), _p(mod), L,
                          ctypes.c_size_t(a.shape[0]), int(b_shared), _p(out))
    if rc:
        raise ValueError("orc_modmul rc=%d" % rc)
    return out


def encrypt(n, hs, pt, r, make_secure=True):
    """n (NL,), hs (2NL,) or None, pt (count, NL), r (count, RL)."""
    n, hs, pt, r = _c(n), _c(hs), _c(pt), _c(r)
    NL = n.shape[-1]
    count = pt.shape[0]
    ct = np.zeros((count, 2 * NL), dtype=np.uint32)
    rc = lib().orc_encrypt(_p(n), NL, _p(hs), _p(pt), _p(r),
                           0 if r is None else r.shape[-1],
                           ctypes.c_size_t(count), int(make_secure), _p(ct))
    if rc:
        raise ValueError("orc_encrypt rc=%d" % rc)
    return ct


def decrypt_crt(p, q, ct):
    p, q, ct = _c(p), _c(q), _c(ct)
    PL = p.shape[-1]
    pt = np.zeros((ct.shape[0], 2 * PL), dtype=np.uint32)
    rc = lib().orc_decrypt_crt(_p(p), _p(q), PL, _p(ct),
                               ctypes.c_size_t(ct.shape[0]), _p(pt))
    if rc:
        raise ValueError("orc_decrypt_crt rc=%d" % rc)
    return pt


def decrypt_raw(n, lam, x, ct):
    n, lam, x, ct = _c(n), _c(lam), _c(x), _c(ct)
    NL = n.shape[-1]
    pt = np.zeros((ct.shape[0], NL), dtype=np.uint32)
    rc = lib().orc_decrypt_raw(_p(n), NL, _p(lam), _p(x), _p(ct),
                               ctypes.c_size_t(ct.shape[0]), _p(pt))
    if rc:
        raise ValueError("orc_decrypt_raw rc=%d" % rc)
    return pt


# ---- 8-lane AVX512-IFMA restatement of mbx_exp_mb8 (ifma_modexp.c) ----------
def have_ifma():
    L = lib()
    return bool(L.orc_have_ifma())


def modexp_mb8(base, exp, mod, shared_base=False, shared_exp=False):
    """shared modulus only (the shape of every call on the Paillier path)"""
    base, exp, mod = _c(base), _c(exp), _c(mod)
    L, EL = mod.shape[-1], exp.shape[-1]
    count = max(np.atleast_2d(base).shape[0], np.atleast_2d(exp).shape[0])
    out = np.zeros((count, L), dtype=np.uint32)
    rc = lib().orc_modexp_mb8(_p(base), ctypes.c_size_t(0 if shared_base else L),
                              _p(exp), ctypes.c_size_t(0 if shared_exp else EL),
                              EL, _p(mod), L, ctypes.c_size_t(count), _p(out))
    if rc:
        raise ValueError("orc_modexp_mb8 rc=%d" % rc)
    return out


def encrypt_mb8(n, hs, pt, r):
    n, hs, pt, r = _c(n), _c(hs), _c(pt), _c(r)
    NL = n.shape[-1]
    ct = np.zeros((pt.shape[0], 2 * NL), dtype=np.uint32)
    rc = lib().orc_encrypt_mb8(_p(n), NL, _p(hs), _p(pt), _p(r), r.shape[-1],
                               ctypes.c_size_t(pt.shape[0]), _p(ct))
    if rc:
        raise ValueError("orc_encrypt_mb8 rc=%d" % rc)
    return ct


def decrypt_crt_mb8(p, q, ct):
    p, q, ct = _c(p), _c(q), _c(ct)
    PL = p.shape[-1]
    pt = np.zeros((ct.shape[0], 2 * PL), dtype=np.uint32)
    rc = lib().orc_decrypt_crt_mb8(_p(p), _p(q), PL, _p(ct),
                                   ctypes.c_size_t(ct.shape[0]), _p(pt))
    if rc:
        raise ValueError("orc_decrypt_crt_mb8 rc=%d" % rc)
    return pt


# ---- OpenSSL BN_mod_exp_mont_consttime (independent second opinion) ---------
def have_openssl():
    return bool(lib().orc_have_openssl())


def modexp_openssl(base, exp, mod, shared_base=False, shared_exp=False):
    base, exp, mod = _c(base), _c(exp), _c(mod)
    L, EL = mod.shape[-1], exp.shape[-1]
    count = max(np.atleast_2d(base).shape[0], np.atleast_2d(exp).shape[0])
    out = np.zeros((count, L), dtype=np.uint32)
    rc = lib().orc_modexp_openssl(_p(base), ctypes.c_size_t(0 if shared_base else L),
                                  _p(exp), ctypes.c_size_t(0 if shared_exp else EL),
                                  EL, _p(mod), L, ctypes.c_size_t(count), _p(out))
    if rc:
        raise ValueError("orc_modexp_openssl rc=%d" % rc)
    return out
