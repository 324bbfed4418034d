from vdpp_b200.distributed.backend import resolve_backend  # noqa: F401
