"""B200-native SVD denoising-step pipeline (hot path of inai17ibar/video-diffusion-pipeline-parallel).

Sub-packages mirror the reference's interface for the path:
  pipeline/     step assignment + stage runner      (reference src/pipeline/*)
  models/       DummyUNet, StableVideoUNet wrapper, NativeUNet (reference src/models/*)
  distributed/  backend selection + process group   (reference src/distributed/*)
  native.py     ctypes binding of csrc/libsvdpp.so (the C ABI declared in include/svdpp.h)
"""
__all__ = ["pipeline", "models", "distributed"]
__version__ = "0.1.0"
