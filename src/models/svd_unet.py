from vdpp_b200.models.svd_unet import StableVideoUNet  # noqa: F401
