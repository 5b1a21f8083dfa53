import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py

    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def gpu_opt_factory():
    """Factory for BatchOptimizer on cuda:0; fails loudly if the CUDA library is missing."""
    from dftpav_b200.optimizer import BatchOptimizer

    made = []

    def make(prm):
        o = BatchOptimizer(prm, device=0)
        made.append(o)
        return o

    yield make
    for o in made:
        o.close()


@pytest.fixture(scope="session")
def gpu_strict_factory():
    """BatchOptimizer over the strict-order VERIFICATION build (dftpav_b200/csrc/libdftpav_b200_strict.so)."""
    from dftpav_b200.optimizer import BatchOptimizer

    made = []

    def make(prm):
        o = BatchOptimizer(prm, device=0, variant="strict")
        assert o.build_flags() & 1, "not the strict build"
        made.append(o)
        return o

    yield make
    for o in made:
        o.close()
