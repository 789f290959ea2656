"""Runs the C++ test binary that re-expresses the reference's gtests
(test/test_cryptography.cpp, test/test_ops.cpp) against the ipcl:: API of this
repository; the binary calls through libipcl.so -> the C ABI -> the kernels."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def binary():
    from pailliercryptolib_b200 import build
    if not os.path.exists(build.CPP_TEST_BIN):
        build.build_cpp_tests()
    return build.CPP_TEST_BIN


@pytest.mark.parametrize("suite", ["CryptoTest", "OperationTest", "SerialTest",
                                   "DeviceResidentTest"])
def test_cpp_suite(binary, suite):
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([binary, suite], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=900, env=env)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:]
    assert " 0 failed" in r.stdout
