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


@pytest.fixture(autouse=True, scope="session")
def _device_sync_before_host_reads():
    """The engine enqueues on its own (non-blocking) stream and most entry points return before the kernels ran; the tests read results
    with Tensor.cpu(), which copies on torch's current stream. Make every such read wait for the whole device first, so a parity check
    never looks at a half-written buffer."""
    import torch
    if not torch.cuda.is_available():
        yield
        return
    orig = torch.Tensor.cpu

    def cpu(self, *a, **kw):
        if self.is_cuda:
            torch.cuda.synchronize(self.device)
        return orig(self, *a, **kw)
    torch.Tensor.cpu = cpu
    yield
    torch.Tensor.cpu = orig
