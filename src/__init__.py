"""Compatibility namespace: the reference's import paths (``src.pipeline``, ``src.models``,
``src.distributed``) resolve to the B200-native package, so code and tests written against
inai17ibar/video-diffusion-pipeline-parallel run unchanged."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import vdpp_b200  # noqa: E402,F401
