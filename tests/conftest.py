import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402

entry.import_package()  # registers the alias `itnn_b200` for `itensornetworksnext.jl_b200/`


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    return entry.import_oracle()


@pytest.fixture(scope="session")
def pkg():
    return entry.import_package()
