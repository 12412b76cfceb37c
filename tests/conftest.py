import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref built from /root/reference (this container only)")


def pytest_collection_modifyitems(config, items):
    from oracle import ref
    skip_ref = pytest.mark.skip(reason="oracle/_ref (compiled reference) not present")
    for item in items:
        if "ref" in item.keywords and not ref.available():
            item.add_marker(skip_ref)
