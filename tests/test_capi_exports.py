"""The C-ABI library loads and exports every symbol include/ipcl_b200.h
declares (no compute without a GPU), and fails loudly instead of falling back."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from pailliercryptolib_b200 import build, capi
    build.build_cuda()
    return capi.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ipcl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ipclb200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from pailliercryptolib_b200 import capi
    names = declared_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(capi.EXPORTS) == names


def test_version_string(lib):
    assert b"sm_100a" in lib.ipclb200_version()


def test_no_cpu_fallback_without_device(lib):
    """Without a CUDA device compute entry points must fail with NO_DEVICE,
    never silently compute on the host."""
    import numpy as np
    from pailliercryptolib_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("a device is present")
    one = np.array([[3] + [0] * 15], dtype=np.uint32)
    with pytest.raises(capi.IpclB200Error) as ei:
        capi.modexp(one, one, one, capi.SHARED_MOD)
    assert ei.value.code == -4


def test_argument_errors_need_no_device(lib):
    import ctypes
    rc = lib.ipclb200_modexp(None, None, None, 16, 16, ctypes.c_size_t(1), 0, None)
    assert rc == -1
    assert b"null" in lib.ipclb200_last_error()
