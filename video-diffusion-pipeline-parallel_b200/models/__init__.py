from .dummy_unet import DummyUNet
from .svd_unet import StableVideoUNet

__all__ = ["DummyUNet", "StableVideoUNet", "NativeUNet"]


def __getattr__(name):  # NativeUNet pulls in the ctypes binding; import it lazily
    if name == "NativeUNet":
        from .native_unet import NativeUNet
        return NativeUNet
    raise AttributeError(name)
