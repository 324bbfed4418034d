from vdpp_b200.models.dummy_unet import DummyUNet  # noqa: F401
