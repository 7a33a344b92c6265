import os as _os

# four hardware work queues per device (must be set before CUDA is initialised; see c-kzg-4844_b200/csrc/api.cu)
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "4")
import os, sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
