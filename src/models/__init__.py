from vdpp_b200.models import DummyUNet, StableVideoUNet  # noqa: F401

__all__ = ["DummyUNet", "StableVideoUNet"]
