from vdpp_b200.pipeline import *  # noqa: F401,F403
from vdpp_b200.pipeline import __all__  # noqa: F401
