import os
import sys
from pathlib import Path

# Two rank contexts that SHARE one GPU (how the multi-rank paths are tested on a one-GPU box) spin on each other from inside
# their kernels.  With CUDA's default lazy module loading the first launch of a kernel may need a context-wide synchronisation,
# which never completes while the peer's kernel spins (CUDA programming guide, "Lazy Loading -- Concurrent Execution"): observed
# as an occasional 8 s spin-wait time-out in the first multi-rank test of a process.  Must be set before CUDA initialises.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
