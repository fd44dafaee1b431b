import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracleapi
    oracleapi.lib()
    return oracleapi


@pytest.fixture(scope="session")
def ref():
    import refapi
    if not refapi.available():
        pytest.skip("oracle/_ref not built (reference sources not mounted)")
    refapi.lib()
    return refapi
