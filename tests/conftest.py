import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a host without a CUDA device (or without the built library) skips the gpu-marked tests instead
    of failing them; `-m gpu` on the B200 box runs them all."""
    import pytest
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        have_gpu = False
    lib = os.path.join(ROOT, "video-diffusion-pipeline-parallel_b200", "csrc", "libsvdpp.so")
    if have_gpu and os.path.exists(lib):
        return
    why = "no CUDA device" if not have_gpu else "csrc/libsvdpp.so not built"
    skip = pytest.mark.skip(reason=f"gpu test: {why}")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
