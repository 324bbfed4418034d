"""GPU parity tests proper: every libsvdpp.so entry point against the same op in plain torch, through the
C ABI (ctypes).  Sizes finish in seconds; full-size properties live in test_gpu_unet.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import kernel_checks as kc
    NAMES = sorted(kc.ALL_CHECKS)
else:
    NAMES = []


def test_device_is_blackwell():
    from vdpp_b200 import native
    major, minor, sms = native.device_info()
    assert major == 10, f"libsvdpp.so is built for sm_100a only, found sm_{major}{minor}"
    assert sms > 0


@pytest.mark.parametrize("name", NAMES)
def test_kernel(name):
    r = kc.ALL_CHECKS[name]()
    assert r["ok"], r
