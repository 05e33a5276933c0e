import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Both shared libraries, built in-tree (no-op when up to date)."""
    from noahmp_b200 import _lib
    from oracle import oracle as O
    _lib.build()
    O.build()
    return True


@pytest.fixture(scope="session")
def tables_usgs():
    from noahmp_b200 import tables
    return tables.default_tables("USGS")


@pytest.fixture(scope="session")
def tables_usgs_struct(tables_usgs):
    from noahmp_b200 import _capi
    return _capi.tables_from_dict(tables_usgs)
