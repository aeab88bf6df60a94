import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.dirname(os.path.abspath(__file__)) not in sys.path:
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "container: needs /root/reference (build container only)")


def _cuda_device_absent():
    """True only when a CUDA runtime was found AND reports no usable device; unknown (no libcudart found) counts as present,
    so that a GPU box can never skip the parity tests by accident."""
    import ctypes
    import glob
    cands = ["libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"] + sorted(glob.glob("/usr/local/cuda*/lib64/libcudart.so*"))
    for name in cands:
        try:
            rt = ctypes.CDLL(name)
        except OSError:
            continue
        n = ctypes.c_int(0)
        try:
            rc = rt.cudaGetDeviceCount(ctypes.byref(n))
        except Exception:
            return False
        return rc != 0 or n.value == 0
    return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a GPU skips the gpu-marked tests instead of failing them on VGL_ENODEV."""
    if not _cuda_device_absent():
        return
    skip = pytest.mark.skip(reason="no CUDA device: libvgl has no CPU path (run with -m gpu on a B200 box)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
