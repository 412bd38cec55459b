"""pytest configuration: registers the `gpu` marker and shared fixtures.

Everything that needs a GPU is marked @pytest.mark.gpu; the rest runs on CPU in a few minutes.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CNR = os.path.join(GOLDEN, "cnr-2000", "cnr-2000")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from tests import oracle_binding
    return oracle_binding.load()


@pytest.fixture(scope="session")
def cnr_truth():
    from tests import oracle_binding
    return oracle_binding.read_ascii_graph(CNR + ".graph-txt.gz")
