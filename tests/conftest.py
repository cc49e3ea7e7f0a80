import gzip
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return json.load(open(os.path.join(GOLDEN_DIR, "golden.json")))


@pytest.fixture(scope="session")
def spneumoniae_bytes():
    return gzip.open(os.path.join(GOLDEN_DIR, "spneumoniae.fa.gz")).read()


@pytest.fixture(scope="session")
def simplitigs_bytes():
    return gzip.open(os.path.join(GOLDEN_DIR, "simplitigs-k31.fa.gz")).read()


@pytest.fixture(scope="session")
def test_fa_bytes():
    return open(os.path.join(GOLDEN_DIR, "test.fa"), "rb").read()


@pytest.fixture(scope="session")
def ctx():
    """One libkcgpu context for the whole GPU session (fails loudly without a device)."""
    import kmercamel_b200 as kb
    c = kb.Context(0)
    yield c
    c.close()


def md5(b: bytes) -> str:
    import hashlib
    return hashlib.md5(b).hexdigest()


def keys_md5(keys: np.ndarray) -> str:
    return md5(np.ascontiguousarray(keys, dtype="<u8").tobytes())
