"""ctypes binding of ``csrc/libsvdpp.so`` — the C ABI declared in ``include/svdpp.h``.

There is no fallback: if the library is missing or a call fails, ``NativeError`` is raised.  Wrappers
take torch CUDA tensors, pass raw device pointers and the current CUDA stream, and return nothing
(outputs are caller-allocated), so every call is CUDA-graph capturable.
"""
from __future__ import annotations

import ctypes as C
import os
import pathlib
from typing import Optional, Sequence, Tuple

import torch

_LIB_PATH = pathlib.Path(__file__).resolve().parent / "csrc" / "libsvdpp.so"
MAX_TAPS = 9
GEMM_BN = 160  # N tile of the tcgen05 GEMM; weights are padded / GEGLU-interleaved to it

EXPORTED_SYMBOLS = (
    "svdpp_abi_version", "svdpp_last_error", "svdpp_device_info", "svdpp_set_tuning", "svdpp_get_tuning",
    "svdpp_gemm_f16", "svdpp_ff_geglu_f16",
    "svdpp_attn_spatial_f16", "svdpp_debug_attn_trace", "svdpp_attn_temporal_f16", "svdpp_groupnorm_workspace_bytes",
    "svdpp_groupnorm_silu", "svdpp_layernorm", "svdpp_linear_small", "svdpp_linear_small_grouped",
    "svdpp_sinusoid_embed",
    "svdpp_upsample2x_nhwc", "svdpp_im2col_nhwc", "svdpp_pack_unet_input", "svdpp_nhwc_to_bfchw",
    "svdpp_euler_vpred_step", "svdpp_euler_vpred_step_signal", "svdpp_flag_wait", "svdpp_flag_set", "svdpp_dummy_unet_step",
    "svdpp_softmax_rows", "svdpp_transpose_f16", "svdpp_time_conv_out", "svdpp_attn_small_f16", "svdpp_frames_to_bytes",
    "svdpp_unet_step_handoff", "svdpp_unet_create", "svdpp_unet_load_weights", "svdpp_unet_weight_bytes", "svdpp_unet_workspace_bytes",
    "svdpp_unet_forward", "svdpp_unet_forward_nhwc", "svdpp_unet_step", "svdpp_unet_last_launches", "svdpp_unet_destroy",
)


class NativeError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64),
        ("A2", C.c_void_p), ("lda2", C.c_int64), ("K1", C.c_int32),
        ("conv", C.c_int32),
        ("cB", C.c_int32), ("cF", C.c_int32), ("cH", C.c_int32), ("cW", C.c_int32), ("cC", C.c_int32),
        ("ntaps", C.c_int32),
        ("taps", (C.c_int8 * 4) * MAX_TAPS),
        ("Wt", C.c_void_p), ("ldw", C.c_int64),
        ("bias", C.c_void_p),
        ("rowvec", C.c_void_p), ("rv_ld", C.c_int64),
        ("rv_hw", C.c_int32), ("rv_div", C.c_int32), ("rv_mod", C.c_int32),
        ("R1", C.c_void_p), ("ldr1", C.c_int64), ("beta1", C.c_float),
        ("R2", C.c_void_p), ("ldr2", C.c_int64), ("beta2", C.c_float),
        ("alpha", C.c_float),
        ("geglu", C.c_int32),
        ("D", C.c_void_p), ("ldd", C.c_int64),
        ("n_store", C.c_int32),
        ("conv_stride", C.c_int32), ("cHin", C.c_int32), ("cWin", C.c_int32),
        ("out_up", C.c_int32), ("out_up_y", C.c_int32), ("out_up_x", C.c_int32),
        ("splitk_ws", C.c_void_p), ("splitk_ws_bytes", C.c_int64),
    ]


class FfDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("C", C.c_int32),
        ("X", C.c_void_p), ("ldx", C.c_int64),
        ("W1", C.c_void_p), ("b1", C.c_void_p),
        ("W2", C.c_void_p), ("ldw2", C.c_int64), ("w2_rows", C.c_int32),
        ("b2", C.c_void_p),
        ("rowvec", C.c_void_p), ("rv_ld", C.c_int64),
        ("rv_hw", C.c_int32), ("rv_div", C.c_int32), ("rv_mod", C.c_int32),
        ("R1", C.c_void_p), ("ldr1", C.c_int64), ("beta1", C.c_float),
        ("R2", C.c_void_p), ("ldr2", C.c_int64), ("beta2", C.c_float),
        ("alpha", C.c_float),
        ("D", C.c_void_p), ("ldd", C.c_int64),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("qkv", C.c_void_p), ("ld", C.c_int64),
        ("q_off", C.c_int32), ("k_off", C.c_int32), ("v_off", C.c_int32),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("n_img", C.c_int32), ("S", C.c_int32), ("heads", C.c_int32),
        ("scale", C.c_float),
    ]


UNET_MAX_LEVELS = 8


class UNetConfig(C.Structure):
    """``svdpp_unet_config`` (include/svdpp.h)."""
    _fields_ = [
        ("in_channels", C.c_int32), ("out_channels", C.c_int32), ("n_levels", C.c_int32),
        ("block_out_channels", C.c_int32 * UNET_MAX_LEVELS), ("down_attn", C.c_int32 * UNET_MAX_LEVELS),
        ("num_attention_heads", C.c_int32 * UNET_MAX_LEVELS),
        ("layers_per_block", C.c_int32), ("cross_attention_dim", C.c_int32), ("addition_time_embed_dim", C.c_int32),
        ("projection_class_embeddings_input_dim", C.c_int32),
        ("eps_down_attn", C.c_float), ("eps_down", C.c_float), ("eps_mid", C.c_float), ("eps_up", C.c_float),
        ("eps_transformer", C.c_float), ("eps_out", C.c_float),
        ("gemm_impl", C.c_int32), ("attn_impl", C.c_int32), ("attn_impl_long", C.c_int32),
    ]


class Handoff(C.Structure):
    """``svdpp_handoff``: completion counter (local), ready flag (in the consumer's memory), value to store."""
    _fields_ = [("done_counter", C.c_void_p), ("ready_flag", C.c_void_p), ("flag_value", C.c_uint32)]


class TensorDesc(C.Structure):
    """``svdpp_tensor_desc``."""
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int32), ("shape", C.c_int64 * 5),
                ("dtype", C.c_int32)]


_lib = None

# Number of libsvdpp kernels launched through this binding (bench.py reports it as gpu_launches).
LAUNCHES = 0
# Optional profiling hook: when set to a list, gemm()/attn_spatial() append (kind, flops, start_evt, end_evt).
PROFILE = None


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def library_path() -> pathlib.Path:
    return _LIB_PATH


def _bind(lib):
    """Declare the argument types of every entry point on a freshly dlopen-ed libsvdpp."""
    lib.svdpp_last_error.restype = C.c_char_p
    lib.svdpp_abi_version.restype = C.c_int
    lib.svdpp_set_tuning.argtypes = [C.c_char_p, C.c_int]
    lib.svdpp_get_tuning.argtypes = [C.c_char_p]
    lib.svdpp_groupnorm_workspace_bytes.restype = C.c_size_t
    lib.svdpp_groupnorm_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
    lib.svdpp_gemm_f16.argtypes = [C.POINTER(GemmDesc), C.c_int, C.c_void_p]
    lib.svdpp_ff_geglu_f16.argtypes = [C.POINTER(FfDesc), C.c_void_p]
    lib.svdpp_attn_spatial_f16.argtypes = [C.POINTER(AttnDesc), C.c_int, C.c_void_p]
    lib.svdpp_debug_attn_trace.argtypes = [C.c_void_p]
    lib.svdpp_attn_temporal_f16.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                            C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                            C.c_void_p]
    lib.svdpp_groupnorm_silu.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32,
                                         C.c_void_p, C.c_size_t, C.c_void_p]
    lib.svdpp_layernorm.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_void_p]
    lib.svdpp_linear_small.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.svdpp_linear_small_grouped.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                               C.c_int64, C.c_int32, C.c_void_p]
    lib.svdpp_sinusoid_embed.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                         C.c_void_p]
    lib.svdpp_upsample2x_nhwc.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_void_p]
    lib.svdpp_im2col_nhwc.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.svdpp_pack_unet_input.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_float,
                                          C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_void_p,
                                          C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.svdpp_nhwc_to_bfchw.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_void_p]
    lib.svdpp_euler_vpred_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float,
                                           C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.svdpp_dummy_unet_step.argtypes = [C.c_void_p] * 7 + [C.c_float, C.c_float, C.c_void_p, C.c_void_p] + \
        [C.c_int32] * 6 + [C.c_void_p]
    lib.svdpp_unet_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(UNetConfig)]
    lib.svdpp_unet_load_weights.argtypes = [C.c_void_p, C.POINTER(TensorDesc), C.c_int]
    lib.svdpp_unet_weight_bytes.restype = C.c_size_t
    lib.svdpp_unet_weight_bytes.argtypes = [C.c_void_p]
    lib.svdpp_unet_workspace_bytes.restype = C.c_size_t
    lib.svdpp_unet_workspace_bytes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    fwd = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
           C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.svdpp_unet_forward.argtypes = fwd
    lib.svdpp_unet_forward_nhwc.argtypes = fwd
    lib.svdpp_unet_step.argtypes = [C.c_void_p] * 7 + [C.c_float] * 6 + [C.c_void_p, C.c_void_p, C.c_size_t,
                                                                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.svdpp_unet_step_handoff.argtypes = lib.svdpp_unet_step.argtypes[:-1] + [C.POINTER(Handoff), C.c_void_p]
    lib.svdpp_euler_vpred_step_signal.argtypes = lib.svdpp_euler_vpred_step.argtypes[:-1] + [C.POINTER(Handoff), C.c_void_p]
    lib.svdpp_flag_wait.argtypes = [C.c_void_p, C.c_uint32, C.c_int32, C.c_uint32, C.c_int32, C.c_void_p]
    lib.svdpp_flag_set.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    lib.svdpp_softmax_rows.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p]
    lib.svdpp_attn_small_f16.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                         C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p]
    lib.svdpp_transpose_f16.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
    lib.svdpp_frames_to_bytes.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    lib.svdpp_time_conv_out.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_int64, C.c_void_p]
    lib.svdpp_unet_last_launches.restype = C.c_longlong
    lib.svdpp_unet_last_launches.argtypes = [C.c_void_p]
    lib.svdpp_unet_destroy.restype = None
    lib.svdpp_unet_destroy.argtypes = [C.c_void_p]
    return lib


def use_library(path) -> None:
    """Swap in another build of libsvdpp.so (A/B measurements of two builds inside one process: tools/ab_lib.py)."""
    global _lib
    lib = _bind(C.CDLL(str(path)))
    if lib.svdpp_abi_version() != 1:
        raise NativeError("libsvdpp.so ABI version mismatch; rebuild")
    _lib = lib


def load():
    """Load libsvdpp.so (once). Raises NativeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise NativeError(
            f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU or library fallback for this path.")
    lib = _bind(C.CDLL(str(_LIB_PATH)))
    if lib.svdpp_abi_version() != 1:
        raise NativeError("libsvdpp.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise NativeError(f"{what} failed ({rc}): {load().svdpp_last_error().decode()}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype=torch.float16) -> None:
    if not t.is_cuda:
        raise NativeError("native kernels need CUDA tensors (there is no CPU path)")
    if t.dtype != dtype:
        raise NativeError(f"expected dtype {dtype}, got {t.dtype}")


def set_tuning(key: str, value: int) -> int:
    """Set a process-wide kernel tuning switch ("tma_store", "pdl"; include/svdpp.h); returns the previous value."""
    old = get_tuning(key)
    _check(load().svdpp_set_tuning(key.encode(), int(value)), "svdpp_set_tuning")
    return old


def get_tuning(key: str) -> int:
    v = load().svdpp_get_tuning(key.encode())
    if v < 0:
        raise NativeError(f"unknown tuning key {key!r}")
    return v


def device_info() -> Tuple[int, int, int]:
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    _check(load().svdpp_device_info(C.byref(a), C.byref(b), C.byref(c)), "svdpp_device_info")
    return a.value, b.value, c.value


# taps ---------------------------------------------------------------------------------------------
TAPS_3X3 = tuple((kw - 1, kh - 1, 0) for kh in range(3) for kw in range(3))       # K order (kh, kw, c)
TAPS_T3 = tuple((0, 0, kt - 1) for kt in range(3))                                # K order (kt, c)
TAPS_1 = ((0, 0, 0),)


def _fill_taps(desc: GemmDesc, taps: Sequence[Tuple[int, int, int]]) -> None:
    desc.ntaps = len(taps)
    for i, (dw, dh, df) in enumerate(taps):
        desc.taps[i][0], desc.taps[i][1], desc.taps[i][2], desc.taps[i][3] = dw, dh, df, 0


SPLITK_WS_BYTES = 4096 + 74 * 2 * 20 * 128 * 16 * 4   # counters + one fp32 partial per CTA pair (include/svdpp.h)
_splitk_ws = {}


def splitk_workspace(device) -> torch.Tensor:
    """Per-device scratch of the split-K tail of impl 6 (zeroed once; the kernel re-arms its counters).  All GEMMs
    of a process run on one stream, so one buffer serves them all.  Create it outside CUDA-graph capture."""
    dev = torch.device(device)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _splitk_ws:
        _splitk_ws[key] = torch.zeros(SPLITK_WS_BYTES, dtype=torch.uint8, device=dev)
    return _splitk_ws[key]


def gemm(out: torch.Tensor, a: torch.Tensor, w: torch.Tensor, *, bias=None, a2=None,
         conv_dims: Optional[Tuple[int, int, int, int, int]] = None, taps=None,
         rowvec=None, rv_hw=1, rv_div=1, rv_mod=0, r1=None, beta1=1.0, r2=None, beta2=1.0, alpha=1.0,
         geglu=False, n_store=0, impl=0, conv_stride: int = 1,
         conv_in_hw: Optional[Tuple[int, int]] = None, out_up: Optional[Tuple[int, int, int]] = None,
         splitk_ws: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``out = epilogue(A @ w.T)``; see ``svdpp_gemm_desc`` in include/svdpp.h.

    ``a``: [M, K] (row stride may exceed K) or, with ``conv_dims=(B,F,H,W,C)``, the contiguous
    channels-last activation.  ``w``: [N, K] fp16, N a multiple of the tile width (GEGLU-interleaved if geglu).
    ``conv_stride=2`` with ``conv_in_hw=(Hin, Win)``: strided windows (conv_dims then holds the OUTPUT H, W).
    ``out_up=(2, py, px)``: row (img, h, w) is stored at pixel (2h+py, 2w+px) of a 2H x 2W output image.
    """
    lib = load()
    d = GemmDesc()
    for t in (out, a, w):
        _req(t)
    N, K = w.shape
    d.N, d.K = N, K
    d.Wt, d.ldw = w.data_ptr(), w.stride(0)
    if conv_dims is not None:
        B_, F_, H_, W_, C_ = conv_dims
        d.conv = 1
        d.cB, d.cF, d.cH, d.cW, d.cC = B_, F_, H_, W_, C_
        _fill_taps(d, taps)
        d.M = B_ * F_ * H_ * W_
        if not a.is_contiguous():
            raise NativeError("conv-mode activation must be contiguous")
        d.A, d.lda = a.data_ptr(), C_
        if conv_stride > 1:
            if conv_in_hw is None:
                raise NativeError("strided conv needs conv_in_hw=(Hin, Win)")
            d.conv_stride, d.cHin, d.cWin = conv_stride, conv_in_hw[0], conv_in_hw[1]
    else:
        d.conv = 0
        d.M = a.shape[0]
        d.A, d.lda = a.data_ptr(), a.stride(0)
        if a2 is not None:
            _req(a2)
            d.A2, d.lda2, d.K1 = a2.data_ptr(), a2.stride(0), a.shape[1]
    d.bias = _ptr(bias)
    if rowvec is not None:
        d.rowvec, d.rv_ld = rowvec.data_ptr(), rowvec.stride(0)
        d.rv_hw, d.rv_div, d.rv_mod = rv_hw, rv_div, rv_mod
    if r1 is not None:
        d.R1, d.ldr1, d.beta1 = r1.data_ptr(), r1.stride(0), beta1
    if r2 is not None:
        d.R2, d.ldr2, d.beta2 = r2.data_ptr(), r2.stride(0), beta2
    d.alpha = alpha
    d.geglu = 1 if geglu else 0
    d.D, d.ldd = out.data_ptr(), out.stride(0)
    if splitk_ws is None and impl == 6 and not torch.cuda.is_current_stream_capturing():
        splitk_ws = splitk_workspace(out.device)
    elif splitk_ws is None and impl == 6:
        splitk_ws = _splitk_ws.get(out.device.index)        # during capture: only a buffer that already exists
    if splitk_ws is not None:
        d.splitk_ws, d.splitk_ws_bytes = splitk_ws.data_ptr(), splitk_ws.numel() * splitk_ws.element_size()
    if out_up is not None:
        d.out_up, d.out_up_y, d.out_up_x = out_up
    d.n_store = n_store
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _check(lib.svdpp_gemm_f16(C.byref(d), impl, _stream()), "svdpp_gemm_f16")
    if PROFILE is not None:
        e1.record()
        kind = "conv" if conv_dims is not None else ("geglu" if geglu else "linear")
        PROFILE.append((kind, 2.0 * d.M * N * K, (d.M, N, K), e0, e1))
    _count()
    return out


def ff_geglu(out: torch.Tensor, x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor, *,
             rowvec=None, rv_hw: int = 1, rv_div: int = 1, rv_mod: int = 0, r1=None, beta1: float = 1.0, r2=None,
             beta2: float = 1.0, alpha: float = 1.0) -> torch.Tensor:
    """Fused feed-forward (svdpp_ff_geglu_f16): ``w1`` / ``b1`` interleaved per 128 rows as [64 value | 64 gate]
    (``interleave_geglu(w, b, half=64)``), ``w2`` [>= C, 4C] as stored by diffusers (rows may be padded)."""
    _req(out), _req(x), _req(w1), _req(w2)
    d = FfDesc()
    d.M, d.C = x.shape[0], x.shape[1]
    d.X, d.ldx = x.data_ptr(), x.stride(0)
    d.W1, d.b1 = w1.data_ptr(), b1.data_ptr()
    d.W2, d.ldw2, d.w2_rows = w2.data_ptr(), w2.stride(0), w2.shape[0]
    d.b2 = b2.data_ptr()
    if rowvec is not None:
        d.rowvec, d.rv_ld = rowvec.data_ptr(), rowvec.stride(0)
    d.rv_hw, d.rv_div, d.rv_mod = rv_hw, rv_div, rv_mod
    if r1 is not None:
        d.R1, d.ldr1 = r1.data_ptr(), r1.stride(0)
    if r2 is not None:
        d.R2, d.ldr2 = r2.data_ptr(), r2.stride(0)
    d.beta1, d.beta2, d.alpha = beta1, beta2, alpha
    d.D, d.ldd = out.data_ptr(), out.stride(0)
    _check(load().svdpp_ff_geglu_f16(C.byref(d), _stream()), "svdpp_ff_geglu_f16")
    _count(1)
    return out


def attn_spatial(out: torch.Tensor, qkv: torch.Tensor, *, n_img: int, S: int, heads: int,
                 q_off: int, k_off: int, v_off: int, scale: float, impl=0) -> torch.Tensor:
    _req(out), _req(qkv)
    d = AttnDesc()
    d.qkv, d.ld = qkv.data_ptr(), qkv.stride(0)
    d.q_off, d.k_off, d.v_off = q_off, k_off, v_off
    d.out, d.ldo = out.data_ptr(), out.stride(0)
    d.n_img, d.S, d.heads, d.scale = n_img, S, heads, scale
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _check(load().svdpp_attn_spatial_f16(C.byref(d), impl, _stream()), "svdpp_attn_spatial_f16")
    if PROFILE is not None:
        e1.record()
        PROFILE.append(("attn_spatial", 4.0 * S * S * 64 * heads * n_img, (n_img, S, heads), e0, e1))
    _count()
    return out


def attn_temporal(out: torch.Tensor, qkv: torch.Tensor, *, B: int, F: int, HW: int, heads: int,
                  q_off: int, k_off: int, v_off: int, scale: float) -> torch.Tensor:
    _req(out), _req(qkv)
    _check(load().svdpp_attn_temporal_f16(qkv.data_ptr(), qkv.stride(0), q_off, k_off, v_off, out.data_ptr(),
                                          out.stride(0), B, F, HW, heads, scale, _stream()),
           "svdpp_attn_temporal_f16")
    _count(1)
    return out


def groupnorm_workspace_bytes(n_img: int, HW: int) -> int:
    return int(load().svdpp_groupnorm_workspace_bytes(n_img, HW))


def groupnorm_silu(out, x1, gamma, beta, *, n_img, HW, eps, silu=True, x2=None, frames_per_stat=1,
                   workspace: torch.Tensor) -> torch.Tensor:
    _req(out), _req(x1)
    C1 = x1.shape[-1]
    C2 = 0 if x2 is None else x2.shape[-1]
    _check(load().svdpp_groupnorm_silu(x1.data_ptr(), C1, _ptr(x2), C2, gamma.data_ptr(), beta.data_ptr(),
                                       out.data_ptr(), n_img, HW, frames_per_stat, eps, 1 if silu else 0,
                                       workspace.data_ptr(), workspace.numel() * workspace.element_size(),
                                       _stream()), "svdpp_groupnorm_silu")
    _count(3)
    return out


def layernorm(out, x, gamma, beta, *, eps=1e-5, addvec=None, add_hw=1, add_mod=1) -> torch.Tensor:
    _req(out), _req(x)
    M, Cc = x.shape
    _check(load().svdpp_layernorm(x.data_ptr(), x.stride(0), _ptr(addvec), add_hw, add_mod, gamma.data_ptr(),
                                  beta.data_ptr(), out.data_ptr(), out.stride(0), M, Cc, eps, _stream()),
           "svdpp_layernorm")
    _count(1)
    return out


def linear_small(out, x, w, bias=None, *, x_add=None, act_in=0, act_out=0) -> torch.Tensor:
    _req(out), _req(x), _req(w)
    R, K = x.shape
    N = w.shape[0]
    if x_add is not None and (x_add.stride(0) != x.stride(0) or x_add.shape != x.shape):
        raise NativeError("linear_small: x_add must match x")
    _check(load().svdpp_linear_small(x.data_ptr(), _ptr(x_add), x.stride(0), w.data_ptr(), w.stride(0), _ptr(bias),
                                     out.data_ptr(), out.stride(0), R, N, K, act_in, act_out, _stream()),
           "svdpp_linear_small")
    _count(1)
    return out


class SmallGroup(C.Structure):
    _fields_ = [("W", C.c_void_p), ("bias", C.c_void_p), ("x_off", C.c_int32), ("y_off", C.c_int32),
                ("N", C.c_int32), ("K", C.c_int32)]


def pack_small_groups(groups, device) -> torch.Tensor:
    """[(W [N,K] contiguous, bias or None, x_off, y_off)] -> device table of ``svdpp_small_group``."""
    arr = (SmallGroup * len(groups))()
    for i, (w, b, x_off, y_off) in enumerate(groups):
        _req(w)
        if not w.is_contiguous() or x_off % 8 or w.shape[1] % 8:
            raise NativeError("small group: W must be contiguous, K and x_off multiples of 8")
        arr[i].W, arr[i].bias = w.data_ptr(), _ptr(b)
        arr[i].x_off, arr[i].y_off, arr[i].N, arr[i].K = x_off, y_off, w.shape[0], w.shape[1]
    raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    return raw.to(device)


def linear_small_grouped(out, x, table: torch.Tensor, *, n_groups: int, max_n: int) -> torch.Tensor:
    """All groups of ``table`` (see ``pack_small_groups``) in one launch; x: [R, *], out: [R, *]."""
    _req(out), _req(x)
    _check(load().svdpp_linear_small_grouped(x.data_ptr(), x.stride(0), table.data_ptr(), n_groups, max_n,
                                             out.data_ptr(), out.stride(0), x.shape[0], _stream()),
           "svdpp_linear_small_grouped")
    _count(1)
    return out


def sinusoid_embed(out, src: Optional[torch.Tensor], *, n_vals: int, dim: int, src_mod: int = 0) -> torch.Tensor:
    _req(out)
    if src is None:
        kind = 2
    elif src.dtype == torch.float32:
        kind = 0
    elif src.dtype == torch.float16:
        kind = 1
    else:
        raise NativeError("sinusoid_embed: src must be fp32 or fp16")
    _check(load().svdpp_sinusoid_embed(_ptr(src), kind, src_mod, n_vals, dim, out.data_ptr(), _stream()),
           "svdpp_sinusoid_embed")
    _count(1)
    return out


def upsample2x(out, x, *, n_img, H, W, Cc) -> torch.Tensor:
    _req(out), _req(x)
    _check(load().svdpp_upsample2x_nhwc(x.data_ptr(), out.data_ptr(), n_img, H, W, Cc, _stream()),
           "svdpp_upsample2x_nhwc")
    _count(1)
    return out


def im2col(out, x, *, B, F, H, W, Cc, Ho, Wo, stride, taps) -> torch.Tensor:
    _req(out), _req(x)
    arr = (C.c_int8 * (4 * len(taps)))()
    for i, (dw, dh, df) in enumerate(taps):
        arr[4 * i], arr[4 * i + 1], arr[4 * i + 2], arr[4 * i + 3] = dw, dh, df, 0
    _check(load().svdpp_im2col_nhwc(x.data_ptr(), out.data_ptr(), out.stride(0), B, F, H, W, Cc, Ho, Wo, stride,
                                    len(taps), C.cast(arr, C.c_void_p), _stream()), "svdpp_im2col_nhwc")
    _count(1)
    return out


def pack_unet_input(out, src0, strides0, C0, in_div, src1, strides1, C1, *, B, F, H, W,
                    out_bfchw: bool = False) -> torch.Tensor:
    _req(out), _req(src0)
    s1 = strides1 if src1 is not None else (0, 0, 0)
    _check(load().svdpp_pack_unet_input(src0.data_ptr(), strides0[0], strides0[1], strides0[2], C0, in_div,
                                        _ptr(src1), s1[0], s1[1], s1[2], C1, out.data_ptr(),
                                        1 if out_bfchw else 0, B, F, H, W, _stream()), "svdpp_pack_unet_input")
    _count(1)
    return out


def nhwc_to_bfchw(out, x, *, B, F, Cc, H, W) -> torch.Tensor:
    _req(out), _req(x)
    _check(load().svdpp_nhwc_to_bfchw(x.data_ptr(), out.data_ptr(), B, F, Cc, H, W, _stream()),
           "svdpp_nhwc_to_bfchw")
    _count(1)
    return out


def make_handoff(handoff) -> Optional["Handoff"]:
    """(done_counter_ptr, ready_flag_ptr, value) -> svdpp_handoff, or None."""
    if handoff is None:
        return None
    h = Handoff()
    h.done_counter, h.ready_flag, h.flag_value = int(handoff[0]), int(handoff[1]), int(handoff[2])
    return h


def euler_vpred_step(out, latent, v_a, *, v_cond=None, gs=None, v_nhwc: bool, c_v: float, c_x: float,
                     sigma: float, dt: float, handoff=None) -> torch.Tensor:
    """``handoff=(done_counter_ptr, ready_flag_ptr, value)``: ``out`` is a peer-mapped receive slot and the kernel
    raises the consumer's flag when its stores are complete (svdpp_euler_vpred_step_signal)."""
    _req(out), _req(latent), _req(v_a)
    B, Cc, F, H, W = latent.shape
    h = make_handoff(handoff)
    _check(load().svdpp_euler_vpred_step_signal(latent.data_ptr(), v_a.data_ptr(), _ptr(v_cond), _ptr(gs),
                                                1 if v_nhwc else 0, c_v, c_x, sigma, dt, out.data_ptr(), B, Cc, F, H, W,
                                                C.byref(h) if h is not None else None, _stream()),
           "svdpp_euler_vpred_step")
    _count(1)
    return out


def softmax_rows(x: torch.Tensor, scale: float = 1.0, n_valid: int = 0) -> torch.Tensor:
    """In-place ``softmax(scale * x, dim=-1)`` of an fp16 matrix (row pitch may exceed the row length); columns
    ``>= n_valid`` (0: none) are padding keys and come out as 0."""
    _req(x)
    _check(load().svdpp_softmax_rows(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], n_valid, scale, _stream()),
           "svdpp_softmax_rows")
    _count(1)
    return x


def attn_small(out: torch.Tensor, qkv: torch.Tensor, *, n_img: int, S: int, S_pad: int, heads: int, head_dim: int, q_off: int,
               k_off: int, v_off: int, head_stride: int, out_head_stride: int, scale: float) -> torch.Tensor:
    """Attention of a short sequence (S <= 512) for every (image, head) in one launch; see ``svdpp_attn_small_f16``."""
    _req(out), _req(qkv)
    if qkv.shape[0] < n_img * S_pad or out.shape[0] < n_img * S_pad:
        raise NativeError("attn_small: qkv / out have fewer than n_img * S_pad rows")
    if max(q_off, k_off, v_off) + (heads - 1) * head_stride + head_dim > qkv.shape[1] or heads * out_head_stride > out.shape[1]:
        raise NativeError("attn_small: head columns exceed the matrix width")
    _check(load().svdpp_attn_small_f16(qkv.data_ptr(), qkv.stride(0), q_off, k_off, v_off, head_stride, out.data_ptr(),
                                       out.stride(0), out_head_stride, n_img, S, S_pad, heads, head_dim, scale, _stream()),
           "svdpp_attn_small_f16")
    _count(1)
    return out


def transpose(out: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """out [C, R] <- x [R, C] (fp16, row pitches taken from the tensors)."""
    _req(out), _req(x)
    _check(load().svdpp_transpose_f16(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), x.shape[0], x.shape[1],
                                      _stream()), "svdpp_transpose_f16")
    _count(1)
    return out


CUBE_LEVELS = (6, 7, 6)       # the fixed colour cube of svdpp_frames_to_bytes: index = r * 42 + g * 6 + b


def cube_palette() -> list:
    """The 252 RGB triples of the cube, flat, padded to 768 entries (PIL ``putpalette``)."""
    lr, lg, lb = CUBE_LEVELS
    pal = [int(round(v)) for r in range(lr) for g in range(lg) for b in range(lb)
           for v in (r * 255 / (lr - 1), g * 255 / (lg - 1), b * 255 / (lb - 1))]
    return pal + [0] * (768 - len(pal))


def frames_to_bytes(frames: torch.Tensor, *, rgb: bool = True, palette: bool = False, dither: bool = True):
    """One decoded video ``[3, F, H, W]`` in [-1, 1] (fp32 / fp16, CUDA; the pixel plane contiguous) -> ``(rgb uint8
    [F, H, W, 3] or None, idx uint8 [F, H, W] or None)``: the reference's ``((x + 1) / 2 * 255).clamp(0, 255).to(uint8)``
    and / or indices into ``cube_palette()``."""
    if not frames.is_cuda or frames.dim() != 4 or frames.shape[0] != 3 or frames.dtype not in (torch.float16, torch.float32):
        raise NativeError("frames_to_bytes: frames must be a CUDA fp16 / fp32 tensor [3, F, H, W]")
    _, F, H, W = frames.shape
    if frames.stride(3) != 1 or frames.stride(2) != W:
        frames = frames.contiguous()
    out_rgb = torch.empty((F, H, W, 3), dtype=torch.uint8, device=frames.device) if rgb else None
    out_idx = torch.empty((F, H, W), dtype=torch.uint8, device=frames.device) if palette else None
    _check(load().svdpp_frames_to_bytes(frames.data_ptr(), 1 if frames.dtype == torch.float32 else 0, frames.stride(0),
                                        frames.stride(1), F, H, W, out_rgb.data_ptr() if rgb else None,
                                        out_idx.data_ptr() if palette else None, 1 if dither else 0, _stream()),
           "svdpp_frames_to_bytes")
    _count(1)
    return out_rgb, out_idx


def time_conv_out(out: torch.Tensor, x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, *, B: int, F: int, HW: int
                  ) -> torch.Tensor:
    """TemporalDecoder.time_conv_out: x channels-last [B*F*HW, >= 3] fp16 -> out [B*F, 3, H, W] (fp16 or fp32)."""
    _req(x), _req(w), _req(bias)
    if out.dtype not in (torch.float16, torch.float32) or not out.is_cuda:
        raise NativeError("time_conv_out: out must be a CUDA fp16 / fp32 tensor")
    _check(load().svdpp_time_conv_out(x.data_ptr(), x.shape[1], w.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                      1 if out.dtype == torch.float32 else 0, B, F, HW, _stream()), "svdpp_time_conv_out")
    _count(1)
    return out


def flag_wait(flag_ptr: int, value: int, *, reset_to: Optional[int] = None, timeout_s: int = 600) -> None:
    """Stream-ordered wait until the uint32 at ``flag_ptr`` (written by a peer GPU) equals ``value``."""
    _check(load().svdpp_flag_wait(flag_ptr, value, 0 if reset_to is None else 1, 0 if reset_to is None else reset_to,
                                  timeout_s, _stream()), "svdpp_flag_wait")
    _count(1)


def flag_set(flag_ptr: int, value: int) -> None:
    """Stream-ordered release store of ``value`` to the uint32 at ``flag_ptr`` (possibly peer-mapped)."""
    _check(load().svdpp_flag_set(flag_ptr, value, _stream()), "svdpp_flag_set")
    _count(1)


def dummy_unet_step(out, x, w1, b1, w2, b2, ln_g, ln_b, ln_eps, tanh_scale, hidden_ws) -> torch.Tensor:
    for t in (out, x, w1, b1, w2, b2, hidden_ws):
        _req(t, torch.float32)
    B, Cc, F, H, W = x.shape
    Ch = w1.shape[0]
    _check(load().svdpp_dummy_unet_step(x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                        _ptr(ln_g), _ptr(ln_b), ln_eps, tanh_scale, hidden_ws.data_ptr(),
                                        out.data_ptr(), B, Cc, Ch, F, H, W, _stream()), "svdpp_dummy_unet_step")
    _count(3)
    return out


# ------------------------------------------------------------------------------------------ whole-UNet handle
class UNetHandle:
    """Owner of one ``svdpp_unet`` (include/svdpp.h): create + load_weights here, forward / step below; destroyed with
    the object.  The activation workspace is a torch tensor the caller keeps (one per handle is enough)."""

    def __init__(self, cfg: dict, state_dict, *, gemm_impl: int = 3, attn_impl: Optional[int] = None,
                 attn_impl_long: int = 0):
        lib = load()
        c = UNetConfig()
        boc = tuple(cfg["block_out_channels"])
        if len(boc) > UNET_MAX_LEVELS:
            raise NativeError(f"at most {UNET_MAX_LEVELS} resolution levels")
        c.in_channels, c.out_channels, c.n_levels = cfg["in_channels"], cfg["out_channels"], len(boc)
        for i, v in enumerate(boc):
            c.block_out_channels[i] = v
            c.down_attn[i] = 1 if tuple(cfg["down_attn"])[i] else 0
            c.num_attention_heads[i] = tuple(cfg["num_attention_heads"])[i]
        c.layers_per_block = cfg["layers_per_block"]
        c.cross_attention_dim = cfg["cross_attention_dim"]
        c.addition_time_embed_dim = cfg["addition_time_embed_dim"]
        c.projection_class_embeddings_input_dim = cfg["projection_class_embeddings_input_dim"]
        eps = cfg["norm_eps"]
        c.eps_down_attn, c.eps_down, c.eps_mid, c.eps_up = eps["down_attn"], eps["down"], eps["mid"], eps["up"]
        c.eps_transformer, c.eps_out = eps["transformer"], eps["out"]
        c.gemm_impl = gemm_impl
        c.attn_impl = -1 if attn_impl is None else attn_impl
        c.attn_impl_long = attn_impl_long
        h = C.c_void_p()
        _check(lib.svdpp_unet_create(C.byref(h), C.byref(c)), "svdpp_unet_create")
        self._h = h
        self._lib = lib
        keep = []                      # fp16 / contiguous copies must outlive the call
        descs = (TensorDesc * len(state_dict))()
        for i, (name, t) in enumerate(state_dict.items()):
            if not t.is_cuda:
                raise NativeError("svdpp_unet_load_weights needs CUDA tensors")
            if name.endswith("mix_factor") and t.dtype == torch.float32:
                t, dt = t.detach().contiguous(), 1
            else:
                t, dt = t.detach().to(torch.float16).contiguous(), 0
            keep.append(t)
            descs[i].name = name.encode()
            descs[i].data = t.data_ptr()
            descs[i].ndim = t.dim()
            for j, d in enumerate(t.shape):
                descs[i].shape[j] = d
            descs[i].dtype = dt
        try:
            _check(lib.svdpp_unet_load_weights(h, descs, len(state_dict)), "svdpp_unet_load_weights")
        except NativeError:
            lib.svdpp_unet_destroy(h)
            self._h = None
            raise
        del keep

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None and self._lib is not None:
            self._lib.svdpp_unet_destroy(h)

    def weight_bytes(self) -> int:
        return int(self._lib.svdpp_unet_weight_bytes(self._h))

    def workspace_bytes(self, B: int, F: int, H: int, W: int) -> int:
        n = int(self._lib.svdpp_unet_workspace_bytes(self._h, B, F, H, W))
        if n == 0:
            raise NativeError(f"svdpp_unet_workspace_bytes failed: {self._lib.svdpp_last_error().decode()}")
        return n

    def _done(self, rc: int, what: str) -> None:
        _check(rc, what)
        _count(int(self._lib.svdpp_unet_last_launches(self._h)))

    def forward(self, out, sample, timestep: float, enc, ids, ws, *, B, F, H, W, nhwc: bool = False) -> None:
        for t in (out, sample, enc, ids):
            _req(t)
        fn = self._lib.svdpp_unet_forward_nhwc if nhwc else self._lib.svdpp_unet_forward
        self._done(fn(self._h, sample.data_ptr(), float(timestep), enc.data_ptr(), ids.data_ptr(), out.data_ptr(),
                      ws.data_ptr(), ws.numel() * ws.element_size(), B, F, H, W, _stream()), "svdpp_unet_forward")

    def step(self, out, latent, image_latents, uncond_image_latents, enc, ids, gs, ws, *, timestep, in_div, c_v, c_x,
             sigma, dt, handoff=None) -> None:
        for t in (out, latent, image_latents, enc, ids):
            _req(t)
        B, _, F, H, W = latent.shape
        h = make_handoff(handoff)
        self._done(self._lib.svdpp_unet_step_handoff(self._h, latent.data_ptr(), image_latents.data_ptr(),
                                                     _ptr(uncond_image_latents), enc.data_ptr(), ids.data_ptr(), _ptr(gs),
                                                     float(timestep), in_div, c_v, c_x, sigma, dt, out.data_ptr(),
                                                     ws.data_ptr(), ws.numel() * ws.element_size(), B, F, H, W,
                                                     C.byref(h) if h is not None else None, _stream()), "svdpp_unet_step")
