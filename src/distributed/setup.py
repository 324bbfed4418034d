from vdpp_b200.distributed.setup import finalize_distributed, init_distributed  # noqa: F401
