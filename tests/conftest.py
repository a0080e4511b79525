"""Shared fixtures.  GPU tests are marked `gpu`; everything else runs on CPU.

The oracle (oracle/) is test infrastructure: it is imported here and in the tests only.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def keys80():
    return O.keygen(O.PARAMS_80, 123)


@pytest.fixture(scope="session")
def octx80(keys80):
    return O.Context(keys80)


@pytest.fixture(scope="session")
def keys128():
    return O.keygen(O.PARAMS_128, 123)


@pytest.fixture(scope="session")
def octx128(keys128):
    return O.Context(keys128)


@pytest.fixture(scope="session")
def keys80_small():
    """80-bit parameter set with a 24-element LWE key: fast enough for the exact O(N^2) route."""
    return O.keygen(O.small_params(O.PARAMS_80, 24), 321)


@pytest.fixture(scope="session")
def octx80_small(keys80_small):
    return O.Context(keys80_small)


@pytest.fixture(scope="session")
def mkkeys2():
    return O.mk_keygen(O.MK_PARAMS[2], 2, 77)


@pytest.fixture(scope="session")
def mkctx2(mkkeys2):
    return O.MKContext(mkkeys2)


def random_torus(rng, *shape):
    return rng.integers(-(2 ** 31), 2 ** 31, size=shape, dtype=np.int64).astype(np.int32)


PLAIN_GATES = {
    O.NAND: lambda x, y: ~(x & y), O.OR: lambda x, y: x | y, O.AND: lambda x, y: x & y,
    O.XOR: lambda x, y: x ^ y, O.XNOR: lambda x, y: ~(x ^ y), O.NOR: lambda x, y: ~(x | y),
    O.ANDNY: lambda x, y: (~x) & y, O.ANDYN: lambda x, y: x & (~y), O.ORNY: lambda x, y: (~x) | y,
    O.ORYN: lambda x, y: x | (~y),
}
