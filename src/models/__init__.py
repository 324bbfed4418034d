"""``src.models`` — the reference's import path, served by the B200 package."""
import vdpp_b200.models as _models

DummyUNet = _models.DummyUNet
StableVideoUNet = _models.StableVideoUNet

__all__ = ("DummyUNet", "StableVideoUNet")
