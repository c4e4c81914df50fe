import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 GPUs on the box")
    config.addinivalue_line("markers", "refonly: needs /root/reference (CPU container only)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
