from vdpp_b200.distributed import *  # noqa: F401,F403
