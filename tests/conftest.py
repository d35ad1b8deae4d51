import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "gpu_next: needs a CUDA device; written after the round's GPU budget was spent and never run on one yet — "
                                       "run with `-m gpu_next` first thing, then promote to `gpu`")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import gem_oracle

    gem_oracle.build()
    return gem_oracle
