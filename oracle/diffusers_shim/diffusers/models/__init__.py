from oracle.unet_torch import UNetSpatioTemporalConditionModel  # noqa: F401
