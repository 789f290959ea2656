import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) device")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def iso():
    return {k: int(v, 16) for k, v in load_golden("iso_18033_6.json").items()}


@pytest.fixture(scope="session")
def keys():
    return {bits: {k: int(v, 16) for k, v in d.items()}
            for bits, d in load_golden("keys.json").items()}


@pytest.fixture(scope="session")
def modexp_vectors():
    return load_golden("modexp_vectors.json")


@pytest.fixture(scope="session")
def scheme_vectors():
    return load_golden("scheme_vectors.json")


@pytest.fixture(scope="session")
def oracle():
    from pailliercryptolib_b200 import build
    build.build_oracle()
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    return orc


@pytest.fixture(scope="session")
def capi():
    """The C-ABI library on a real device (gpu tests only)."""
    from pailliercryptolib_b200 import capi as c
    c.init()
    return c
