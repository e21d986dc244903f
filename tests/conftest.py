import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10
    except Exception:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    """GPU tests never silently pass on a CPU box: without an sm_100 device they are reported as skipped."""
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no sm_100 GPU in this environment (there is no CPU fallback to test instead)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
