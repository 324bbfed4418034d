"""Import bootstrap: exposes the package directory ``video-diffusion-pipeline-parallel_b200/``
(not a valid Python identifier) under the importable name ``vdpp_b200``.

``import vdpp_b200`` from the repo root (tests, bench.py, __graft_entry__.py) resolves to this file,
which loads the real package and replaces itself in ``sys.modules``.
"""
import importlib.util
import pathlib
import sys

_PKG_DIR = pathlib.Path(__file__).resolve().parent / "video-diffusion-pipeline-parallel_b200"


def _load():
    spec = importlib.util.spec_from_file_location(
        "vdpp_b200", _PKG_DIR / "__init__.py", submodule_search_locations=[str(_PKG_DIR)]
    )
    mod = importlib.util.module_from_spec(spec)
    sys.modules["vdpp_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


_load()
