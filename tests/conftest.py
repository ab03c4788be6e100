import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# run last: the GPU tests whose newest checks were written in a session without GPU access (proven there on the CPU
# stand-in / a dry run only), so that with `-x` a surprise in them cannot hide the long-standing parity tests
_LAST = ("tests/test_gpu_vs_reference.py",
         "tests/test_adapters.py::test_scan_renderer_adapters_on_the_device", "tests/test_adapters.py::test_map_adapters_on_the_device",
         "tests/test_adapters.py::test_filter_adapters_on_the_device",
         "tests/test_adapters.py::test_map_adapters_gathers_at_half_resolution_on_the_device")


def _rank(item):
    for k, prefix in enumerate(_LAST):
        if item.nodeid.startswith(prefix):
            return k + 1
    return 0


def pytest_collection_modifyitems(config, items):
    items.sort(key=_rank)                                                      # stable: everything else keeps its order
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
