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


# ---------------------------------------------------------------------------
# the reference's OWN sources, unmodified: /root/reference/test/*.cpp and
# /root/reference/example/*.cpp compiled by build.build_reference_programs()
# (in the container that has /root/reference) against this repo's ipcl::
# headers and libipcl.so; the binaries travel to the GPU box
# ---------------------------------------------------------------------------
def _ref_bin(path):
    from pailliercryptolib_b200 import build
    build.build_reference_programs()
    if not os.path.exists(path):
        pytest.skip("%s was not prebuilt (needs /root/reference at build time)" % path)
    return path


@pytest.mark.gpu
@pytest.mark.parametrize("suite", ["CryptoTest", "OperationTest", "SerialTest"])
def test_reference_unit_tests_unmodified(suite):
    from pailliercryptolib_b200 import build
    binary = _ref_bin(build.REF_UNIT_BIN)
    r = subprocess.run([binary, "--gtest_filter=%s.*" % suite], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=1500)
    tail = r.stdout[-3000:]
    assert r.returncode == 0, tail
    assert "[  FAILED  ]" not in r.stdout, tail
    ran = [l for l in r.stdout.splitlines() if l.startswith("[       OK ]")]
    assert len(ran) >= {"CryptoTest": 3, "OperationTest": 12, "SerialTest": 4}[suite], tail


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["example_encrypt_decrypt", "example_add_mul",
                                  "example_hybridmode"])
def test_reference_examples_unmodified(name):
    from pailliercryptolib_b200 import build
    binary = _ref_bin(build.ref_example_bin(name))
    r = subprocess.run([binary], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:]
    # the examples print "... -- pass" / "... -- fail" per check
    assert "-- fail" not in r.stdout, r.stdout[-3000:]
    if name != "example_hybridmode":
        assert "-- pass" in r.stdout, r.stdout[-3000:]
