"""The C-ABI library loads without a GPU and exports every symbol include/svdpp.h declares."""
import ctypes
import os
import re

import pytest

from vdpp_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "svdpp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(svdpp_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    if not native.library_path().exists():
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(str(native.library_path()))
    names = _declared()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in svdpp.h but not exported"
    assert set(names) == set(native.EXPORTED_SYMBOLS)


def test_loader_and_error_string():
    lib = native.load()
    assert lib.svdpp_abi_version() == 1
    assert isinstance(lib.svdpp_last_error(), bytes)
    assert native.groupnorm_workspace_bytes(25, 9216) == (25 * 1153 * 32 + 25 * 32) * 8 + 2 * 4096 * 4 + 25 * 32 * 16


def test_struct_layout_matches_header():
    # svdpp_gemm_desc / svdpp_attn_desc are passed by pointer: sizes must agree with the C compiler
    import subprocess
    import tempfile
    src = ('#include <stdio.h>\n#include "svdpp.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(svdpp_gemm_desc), '
           'sizeof(svdpp_attn_desc), sizeof(svdpp_small_group), sizeof(svdpp_unet_config), sizeof(svdpp_tensor_desc), '
           'sizeof(svdpp_handoff), sizeof(svdpp_ff_desc));return 0;}\n')
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        a, b, c_, u, t, h, f = map(int, subprocess.check_output([exe]).split())
    assert a == ctypes.sizeof(native.GemmDesc) and b == ctypes.sizeof(native.AttnDesc)
    assert c_ == ctypes.sizeof(native.SmallGroup) == 32
    assert u == ctypes.sizeof(native.UNetConfig) and t == ctypes.sizeof(native.TensorDesc) and h == ctypes.sizeof(native.Handoff)
    assert f == ctypes.sizeof(native.FfDesc)


def test_unet_handle_host_side_errors():
    """svdpp_unet_create validates the architecture on the host (no GPU needed); a missing device is an error, not a crash."""
    lib = native.load()
    cfg = native.UNetConfig()
    h = ctypes.c_void_p()
    assert lib.svdpp_unet_create(ctypes.byref(h), ctypes.byref(cfg)) != 0          # n_levels = 0
    assert b"n_levels" in lib.svdpp_last_error()
    cfg.in_channels, cfg.out_channels, cfg.n_levels, cfg.layers_per_block = 8, 4, 1, 2
    cfg.block_out_channels[0], cfg.down_attn[0], cfg.num_attention_heads[0] = 100, 1, 5
    assert lib.svdpp_unet_create(ctypes.byref(h), ctypes.byref(cfg)) != 0          # channels not a multiple of 64
    assert b"multiple of 64" in lib.svdpp_last_error()
    assert lib.svdpp_unet_workspace_bytes(None, 1, 25, 72, 128) == 0               # no handle
    lib.svdpp_unet_destroy(None)                                                    # tolerated


def test_no_cpu_path():
    import torch
    with pytest.raises(native.NativeError):
        native.layernorm(torch.empty(2, 8, dtype=torch.float16), torch.empty(2, 8, dtype=torch.float16),
                         torch.empty(8, dtype=torch.float16), torch.empty(8, dtype=torch.float16))


def test_tuning_switches_roundtrip():
    # svdpp_set_tuning / svdpp_get_tuning are host-only: defaults, round trip, unknown keys
    assert native.get_tuning("tma_store") == 1 and native.get_tuning("tma_r1") == 1
    assert native.get_tuning("pdl") == 0 and native.get_tuning("epi_dma") == 1
    assert native.get_tuning("zigzag") == 0 and native.get_tuning("reverse") == 0 and native.get_tuning("fmha_stagger") == 900
    old = native.set_tuning("pdl", 1)
    try:
        assert old == 0 and native.get_tuning("pdl") == 1
    finally:
        native.set_tuning("pdl", old)
    with pytest.raises(native.NativeError):
        native.get_tuning("no_such_switch")
    with pytest.raises(native.NativeError):
        native.set_tuning("no_such_switch", 1)
