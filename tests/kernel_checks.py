"""Kernel-level parity checks: each function runs one libsvdpp.so entry point on the GPU and compares it
with the same op written in plain torch (fp32 math on the same fp16 inputs).  Used by the pytest GPU
tests and by tools/gpu_check.py (which runs every check in its own process and logs the numbers).

Each check returns a dict with at least {"max_err", "tol", "ok"}.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

import vdpp_b200  # noqa: F401
from vdpp_b200 import native
from vdpp_b200.models.native_unet import interleave_geglu, subpixel_taps, subpixel_weight

DEV = "cuda"


def _rand(*shape, scale=1.0, seed=None):
    g = torch.Generator(device=DEV)
    g.manual_seed(0 if seed is None else seed)
    return (torch.randn(*shape, device=DEV, generator=g) * scale).half()


def _cmp(got: torch.Tensor, ref: torch.Tensor, rel=2e-3, floor=1e-3):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs().max().item()
    tol = rel * ref.abs().max().item() + floor
    cos = F.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
    finite = bool(torch.isfinite(got).all())
    return dict(max_err=err, tol=tol, cos=cos, ref_absmax=ref.abs().max().item(), ok=bool(finite and err <= tol))


class _Guarded:
    """Output buffer with canary rows before and after it: a kernel that writes outside its output (a bad scatter
    map, an unclipped TMA store, a tail tile) trips ``intact()``.  compute-sanitizer is closed on this pool."""
    PAD, CANARY = 64, 12345.0

    def __init__(self, rows: int, cols: int, dtype=torch.float16):
        self.buf = torch.full((rows + 2 * self.PAD, cols), self.CANARY, device=DEV, dtype=dtype)
        self.out = self.buf[self.PAD:self.PAD + rows]
        self.out.fill_(float("nan"))

    def intact(self) -> bool:
        torch.cuda.synchronize()
        return bool((self.buf[:self.PAD] == self.CANARY).all() and (self.buf[-self.PAD:] == self.CANARY).all())


def _with_guard(res: dict, g: "_Guarded") -> dict:
    res["guard_intact"] = g.intact()
    res["ok"] = bool(res["ok"] and res["guard_intact"])
    return res


def _bn(impl):
    return {3: 256, 5: 256, 4: 128, 7: 128, 6: 320}.get(impl, native.GEMM_BN)


def _pad_n(w, mult=native.GEMM_BN):
    n = w.shape[0]
    npad = (n + mult - 1) // mult * mult
    if npad == n:
        return w.contiguous()
    out = torch.zeros((npad,) + tuple(w.shape[1:]), dtype=w.dtype, device=w.device)
    out[:n] = w
    return out


# ------------------------------------------------------------------------------------------ GEMM
def gemm_linear(M=300, N=320, K=320, impl=0, epilogue="full", split=False):
    a = _rand(M, K, seed=1)
    w = _rand(N, K, scale=K ** -0.5, seed=2)
    bias = _rand(N, seed=3)
    guard = _Guarded(M, N)
    out = guard.out
    kw = dict(bias=_pad_n(bias, _bn(impl)))
    ref = a.float() @ w.float().t() + bias.float()
    if epilogue == "full":
        hw, div, mod = 7, 2, 5
        rv = _rand(mod, N, seed=4)
        r1, r2 = _rand(M, N, seed=5), _rand(M, N, seed=6)
        rows = (torch.arange(M, device=DEV) // hw // div) % mod
        ref = 0.75 * (ref + rv.float()[rows]) + 0.5 * r1.float() - 1.25 * r2.float()
        kw.update(rowvec=rv, rv_hw=hw, rv_div=div, rv_mod=mod, r1=r1, beta1=0.5, r2=r2, beta2=-1.25, alpha=0.75)
    if split:
        k1 = (K // 2) // 64 * 64
        a1, a2 = a[:, :k1].contiguous(), a[:, k1:].contiguous()
        native.gemm(out, a1, _pad_n(w, _bn(impl)), a2=a2, n_store=N, impl=impl, **kw)
    else:
        native.gemm(out, a, _pad_n(w, _bn(impl)), n_store=N, impl=impl, **kw)
    torch.cuda.synchronize()
    return _with_guard(_cmp(out, ref), guard)


def gemm_geglu(M=260, C=128, impl=0):
    inner = 4 * C
    a = _rand(M, C, seed=1)
    w = _rand(2 * inner, C, scale=C ** -0.5, seed=2)
    b = _rand(2 * inner, scale=0.1, seed=3)
    wi, bi, n = interleave_geglu(w, b, half=128 if impl == 3 else 80)
    guard = _Guarded(M, inner)
    out = guard.out
    native.gemm(out, a, wi, bias=bi, geglu=True, n_store=inner, impl=impl)
    y = (a.float() @ w.float().t() + b.float()).half()
    val, gate = y.chunk(2, dim=-1)
    ref = val * F.gelu(gate)
    torch.cuda.synchronize()
    return _with_guard(_cmp(out, ref, rel=4e-3), guard)


def conv3x3(B=1, Fr=2, H=6, W=32, C=64, Cout=96, impl=0):
    x = _rand(B * Fr, H, W, C, seed=1)                      # channels-last
    w = _rand(Cout, C, 3, 3, scale=(9 * C) ** -0.5, seed=2)
    b = _rand(Cout, seed=3)
    wk = _pad_n(w.permute(0, 2, 3, 1).reshape(Cout, -1), _bn(impl))
    out = torch.full((B * Fr * H * W, Cout), float("nan"), device=DEV, dtype=torch.float16)
    native.gemm(out, x.reshape(-1, C), wk, bias=_pad_n(b, _bn(impl)), conv_dims=(B, Fr, H, W, C), taps=native.TAPS_3X3,
                n_store=Cout, impl=impl)
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), b.float(), padding=1).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    return _cmp(out, ref.reshape(-1, Cout))


def conv3x3_stride2(B=1, Fr=2, H=12, W=64, C=64, Cout=96, impl=0):
    """Down-sampling Conv2d 3x3 stride 2 pad 1 through the strided-window path (no im2col)."""
    x = _rand(B * Fr, H, W, C, seed=1)
    w = _rand(Cout, C, 3, 3, scale=(9 * C) ** -0.5, seed=2)
    b = _rand(Cout, seed=3)
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    wk = _pad_n(w.permute(0, 2, 3, 1).reshape(Cout, -1), _bn(impl))
    guard = _Guarded(B * Fr * Ho * Wo, Cout)
    out = guard.out
    native.gemm(out, x.reshape(-1, C), wk, bias=_pad_n(b, _bn(impl)), conv_dims=(B, Fr, Ho, Wo, C), taps=native.TAPS_3X3,
                n_store=Cout, impl=impl, conv_stride=2, conv_in_hw=(H, W))
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), b.float(), stride=2, padding=1).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    return _with_guard(_cmp(out, ref.reshape(-1, Cout)), guard)


def conv_up2x(B=1, Fr=2, H=6, W=32, C=64, Cout=96, impl=0):
    """Nearest 2x upsample + Conv2d 3x3 as four parity convolutions with the sub-pixel output map."""
    x = _rand(B * Fr, H, W, C, seed=1)
    w = _rand(Cout, C, 3, 3, scale=(9 * C) ** -0.5, seed=2)
    b = _rand(Cout, seed=3)
    guard = _Guarded(B * Fr * 4 * H * W, Cout)
    out = guard.out
    for py in (0, 1):
        for px in (0, 1):
            wk = _pad_n(subpixel_weight(w, py, px), _bn(impl))
            native.gemm(out, x.reshape(-1, C), wk, bias=_pad_n(b, _bn(impl)), conv_dims=(B, Fr, H, W, C),
                        taps=subpixel_taps(py, px), n_store=Cout, impl=impl, out_up=(2, py, px))
    up = F.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up, w.float(), b.float(), padding=1).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    return _with_guard(_cmp(out, ref.reshape(-1, Cout)), guard)


def conv_temporal(B=2, Fr=5, H=4, W=32, C=64, impl=0):
    x = _rand(B, Fr, H, W, C, seed=1)
    w = _rand(C, C, 3, 1, 1, scale=(3 * C) ** -0.5, seed=2)
    b = _rand(C, seed=3)
    wk = _pad_n(w[:, :, :, 0, 0].permute(0, 2, 1).reshape(C, -1), _bn(impl))
    out = torch.full((B * Fr * H * W, C), float("nan"), device=DEV, dtype=torch.float16)
    native.gemm(out, x.reshape(-1, C), wk, bias=_pad_n(b, _bn(impl)), conv_dims=(B, Fr, H, W, C), taps=native.TAPS_T3,
                n_store=C, impl=impl)
    ref = F.conv3d(x.permute(0, 4, 1, 2, 3).float(), w.float(), b.float(), padding=(1, 0, 0)).permute(0, 2, 3, 4, 1)
    torch.cuda.synchronize()
    return _cmp(out, ref.reshape(-1, C))


def ff_fused(M=1000, C=320, epilogue="res", seed=0):
    """Fused GEGLU feed-forward against torch (fp32 maths on the fp16 weights) and against the two-kernel native path."""
    from vdpp_b200.models.native_unet import interleave_geglu
    inner = 4 * C
    x = _rand(M, C, seed=seed + 1)
    w1 = (_rand(2 * inner, C, seed=seed + 2).float() * C ** -0.5).half()
    b1 = (_rand(2 * inner, seed=seed + 3).float() * 0.1).half()
    w2 = (_rand(C, inner, seed=seed + 4).float() * inner ** -0.5).half()
    b2 = (_rand(C, seed=seed + 5).float() * 0.1).half()
    kw, ref_epi = {}, None
    r1 = r2 = rv = None
    if epilogue in ("res", "blend", "rowvec"):
        r1 = _rand(M, C, seed=seed + 6)
    if epilogue == "blend":
        r2 = _rand(M, C, seed=seed + 7)
        kw = dict(alpha=0.3, beta1=0.3, beta2=0.7)
    if epilogue == "rowvec":
        rv = _rand(5, C, seed=seed + 8)
        kw = dict(rv_hw=7, rv_div=1, rv_mod=5)
    w1i, b1i, _ = interleave_geglu(w1, b1, half=64)
    guard = _Guarded(M, C)
    native.ff_geglu(guard.out, x, w1i, b1i, w2, b2, r1=r1, r2=r2, rowvec=rv, **kw)
    # reference with the fp16 roundings of the two-kernel path
    proj = (x.float() @ w1.float().t() + b1.float()).half()
    h = (proj[:, :inner] * F.gelu(proj[:, inner:].float()).half())
    y = h.float() @ w2.float().t() + b2.float()
    if rv is not None:
        rows = (torch.arange(M, device=DEV) // 7) % 5
        y = y + rv.float()[rows]
    y = kw.get("alpha", 1.0) * y
    if r1 is not None:
        y = y + kw.get("beta1", 1.0) * r1.float()
    if r2 is not None:
        y = y + kw.get("beta2", 1.0) * r2.float()
    torch.cuda.synchronize()
    return _with_guard(_cmp(guard.out, y, rel=4e-3), guard)


# ------------------------------------------------------------------------------------------ attention
def attn_spatial(n_img=2, S=320, heads=2, impl=0, growing=False):
    C = heads * 64
    qkv = _rand(n_img * S, 3 * C, seed=1)
    if growing:
        # keys grow along the sequence so later key blocks raise the row maximum by far more than 2^8:
        # exercises the lazy O-rescale path of the tcgen05 kernel ("late": a jump inside a 128-key block, at key
        # 1000 = quarter 3 of block 7, so quarters of P already stored for the pending PV product get rescaled too)
        pos = (torch.arange(n_img * S, device=DEV) % S).float()
        ramp = ((1.0 + 6.0 * pos / S) if growing is True else torch.where(pos >= 1000, 6.0, 1.0))[:, None]
        qkv = qkv.float()
        qkv[:, :C] *= 2.0
        qkv[:, C:2 * C] *= ramp
        qkv = qkv.half()
    guard = _Guarded(n_img * S, C)
    out = guard.out
    native.attn_spatial(out, qkv, n_img=n_img, S=S, heads=heads, q_off=0, k_off=C, v_off=2 * C, scale=0.125,
                        impl=impl)
    q, k, v = [t.reshape(n_img, S, heads, 64).transpose(1, 2).float() for t in qkv.split(C, dim=1)]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(n_img * S, C)
    torch.cuda.synchronize()
    return _with_guard(_cmp(out, ref, rel=4e-3), guard)


def attn_temporal(B=2, Fr=5, HW=24, heads=2):
    C = heads * 64
    qkv = _rand(B * Fr * HW, 3 * C, seed=1)
    guard = _Guarded(B * Fr * HW, C)
    out = guard.out
    native.attn_temporal(out, qkv, B=B, F=Fr, HW=HW, heads=heads, q_off=0, k_off=C, v_off=2 * C, scale=0.125)
    t = qkv.reshape(B, Fr, HW, 3, heads, 64).permute(3, 0, 2, 4, 1, 5).float()   # [3, B, HW, heads, F, 64]
    ref = F.scaled_dot_product_attention(t[0], t[1], t[2])                       # [B, HW, heads, F, 64]
    ref = ref.permute(0, 3, 1, 2, 4).reshape(B * Fr * HW, C)
    torch.cuda.synchronize()
    return _with_guard(_cmp(out, ref, rel=4e-3), guard)


# ------------------------------------------------------------------------------------------ norms
def groupnorm(n_img=4, HW=100, C1=64, C2=0, fps=1, silu=True, eps=1e-5, in_scale=1.0):
    x1 = ((_rand(n_img * HW, C1, seed=1) * 2 + 0.5).float() * in_scale).half()
    x2 = _rand(n_img * HW, C2, seed=2) if C2 else None
    C = C1 + C2
    g, b = _rand(C, seed=3), _rand(C, seed=4)
    guard = _Guarded(n_img * HW, C)
    out = guard.out
    ws = torch.zeros(native.groupnorm_workspace_bytes(n_img, HW) // 4 + 1, dtype=torch.float32, device=DEV)
    for _ in range(2):   # twice through the same workspace: the arrival counters must come back to zero
        out.fill_(float("nan"))
        native.groupnorm_silu(out, x1, g, b, n_img=n_img, HW=HW, eps=eps, silu=silu, x2=x2, frames_per_stat=fps,
                              workspace=ws)
    x = x1 if x2 is None else torch.cat([x1, x2], dim=1)
    xr = x.float().reshape(n_img // fps, fps * HW, C).permute(0, 2, 1)     # [stat, C, L]
    ref = F.group_norm(xr, 32, g.float(), b.float(), eps)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 1).reshape(n_img * HW, C)
    torch.cuda.synchronize()
    return _with_guard(_cmp(out, ref), guard)


def layernorm(M=70, C=320, add=True):
    x = _rand(M, C, seed=1) * 3 + 1
    g, b = _rand(C, seed=2), _rand(C, seed=3)
    hw, mod = 4, 5
    av = _rand(mod, C, seed=4) if add else None
    out = torch.full((M, C), float("nan"), device=DEV, dtype=torch.float16)
    native.layernorm(out, x, g, b, addvec=av, add_hw=hw, add_mod=mod)
    xin = x
    if add:
        rows = (torch.arange(M, device=DEV) // hw) % mod
        xin = x + av[rows]
    ref = F.layer_norm(xin.float(), (C,), g.float(), b.float(), 1e-5)
    torch.cuda.synchronize()
    return _cmp(out, ref)


def linear_small(R=3, N=100, K=256):
    x, xa = _rand(R, K, seed=1), _rand(R, K, seed=2)
    w, b = _rand(N, K, scale=K ** -0.5, seed=3), _rand(N, seed=4)
    out = torch.full((R, N), float("nan"), device=DEV, dtype=torch.float16)
    native.linear_small(out, x, w, b, x_add=xa, act_in=1, act_out=1)
    ref = F.silu(F.linear(F.silu(x + xa).float(), w.float(), b.float()).half().float())
    torch.cuda.synchronize()
    return _cmp(out, ref)


def linear_small_grouped(R=2, dims=(320, 640, 64, 1280)):
    """Several independent y_g = x[:, off_g:off_g+C_g] @ W_g.T + b_g in one launch."""
    total = sum(dims)
    x = _rand(R, total, seed=1)
    out = torch.full((R, total), float("nan"), device=DEV, dtype=torch.float16)
    groups, ref, off = [], [], 0
    for i, c in enumerate(dims):
        w, b = _rand(c, c, scale=c ** -0.5, seed=10 + i), (_rand(c, seed=30 + i) if i % 2 == 0 else None)
        groups.append((w, b, off, off))
        ref.append(F.linear(x[:, off:off + c].float(), w.float(), None if b is None else b.float()))
        off += c
    table = native.pack_small_groups(groups, DEV)
    native.linear_small_grouped(out, x, table, n_groups=len(dims), max_n=max(dims))
    torch.cuda.synchronize()
    return _cmp(out, torch.cat(ref, dim=1))


def sinusoid(dim=320):
    from oracle.unet_torch import timestep_embedding
    t = torch.tensor([1.6377, -1.5536, 0.3], device=DEV)
    o0 = native.sinusoid_embed(torch.empty(3, dim, device=DEV, dtype=torch.float16), t, n_vals=3, dim=dim)
    r0 = timestep_embedding(t, dim)
    ids = torch.tensor([5.0, 127.0, 0.02], device=DEV).half()
    o1 = native.sinusoid_embed(torch.empty(3, 256, device=DEV, dtype=torch.float16), ids, n_vals=3, dim=256)
    r1 = timestep_embedding(ids, 256)
    o2 = native.sinusoid_embed(torch.empty(6, dim, device=DEV, dtype=torch.float16), None, n_vals=6, dim=dim, src_mod=3)
    r2 = timestep_embedding(torch.arange(3, device=DEV).repeat(2), dim)
    torch.cuda.synchronize()
    res = [_cmp(o0, r0, rel=1e-3), _cmp(o1, r1, rel=1e-3), _cmp(o2, r2, rel=1e-3)]
    return dict(max_err=max(r["max_err"] for r in res), tol=res[0]["tol"], ok=all(r["ok"] for r in res))


# ------------------------------------------------------------------------------------------ data movement
def movers():
    B, Fr, H, W, C = 2, 3, 6, 10, 16
    x = _rand(B * Fr, H, W, C, seed=1)
    up = native.upsample2x(torch.empty(B * Fr, 2 * H, 2 * W, C, device=DEV, dtype=torch.float16), x, n_img=B * Fr,
                           H=H, W=W, Cc=C)
    ref_up = F.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    parts = dict(upsample=torch.equal(up.float(), ref_up))
    ok = True
    for stride in (1, 2):
        Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
        ldo = 9 * C + 16
        cols = torch.full((B * Fr * Ho * Wo, ldo), float("nan"), device=DEV, dtype=torch.float16)
        native.im2col(cols, x, B=B, F=Fr, H=H, W=W, Cc=C, Ho=Ho, Wo=Wo, stride=stride, taps=native.TAPS_3X3)
        unf = F.unfold(x.permute(0, 3, 1, 2).float(), 3, padding=1, stride=stride)      # [n, C*9, L], (c, kh, kw)
        unf = unf.reshape(B * Fr, C, 9, Ho * Wo).permute(0, 3, 2, 1).reshape(-1, 9 * C)  # (kh, kw, c)
        parts[f"im2col_s{stride}"] = torch.equal(cols[:, :9 * C].float(), unf) and bool((cols[:, 9 * C:] == 0).all())
    lat = _rand(B, 4, Fr, H, W, seed=2) * 50
    img = _rand(B, 4, Fr, H, W, seed=3)
    st = (4 * Fr * H * W, H * W, Fr * H * W)
    div = 3.3
    for bfchw in (False, True):
        shape = (B, Fr, 8, H, W) if bfchw else (B * Fr, H, W, 8)
        o = native.pack_unet_input(torch.empty(shape, device=DEV, dtype=torch.float16), lat, st, 4, div, img, st, 4,
                                   B=B, F=Fr, H=H, W=W, out_bfchw=bfchw)
        ref = torch.cat([(lat.float() / div).half(), img], dim=1).permute(0, 2, 1, 3, 4)   # [B,F,8,H,W]
        if not bfchw:
            ref = ref.permute(0, 1, 3, 4, 2).reshape(B * Fr, H, W, 8)
        parts[f"pack_bfchw{int(bfchw)}"] = torch.equal(o, ref.half())
        parts[f"pack_bfchw{int(bfchw)}_err"] = (o.float() - ref.float()).abs().max().item()
    v = _rand(B * Fr * H * W, 4, seed=4)
    o = native.nhwc_to_bfchw(torch.empty(B, Fr, 4, H, W, device=DEV, dtype=torch.float16), v, B=B, F=Fr, Cc=4, H=H, W=W)
    parts["nhwc_to_bfchw"] = torch.equal(o, v.reshape(B, Fr, H, W, 4).permute(0, 1, 4, 2, 3))
    torch.cuda.synchronize()
    ok = all(bool(v_) for k_, v_ in parts.items() if not k_.endswith("_err"))
    return dict(max_err=0.0 if ok else 1.0, tol=0.0, ok=bool(ok), **parts)


def euler(cfg=True, step=3, n=25):
    """Bit-exact against the oracle's restatement of svd_unet.py:410-439 (given the same UNet output)."""
    from oracle.scheduler import euler_karras_tables
    from vdpp_b200.models import scheduler as sched
    B, Fr, H, W = 1, 5, 8, 12
    sig, _, _ = euler_karras_tables(n)
    lat = _rand(B, 4, Fr, H, W, seed=1) * float(sig[step])
    u, c = _rand(B, Fr, 4, H, W, seed=2), _rand(B, Fr, 4, H, W, seed=3)
    gs = torch.linspace(1.0, 3.0, Fr).view(1, 1, Fr, 1, 1).to(DEV, dtype=torch.float16)
    sigma, sigma_next = sig[step].to(DEV), sig[step + 1].to(DEV)
    v = u + gs.permute(0, 2, 1, 3, 4) * (c - u) if cfg else u
    vf = v.permute(0, 2, 1, 3, 4).float()
    x = lat.float()
    s = sigma.float()
    x0 = vf * (-s / (s ** 2 + 1) ** 0.5) + x / (s ** 2 + 1)
    d = (x - x0) / s
    ref = (x + d * (float(sigma_next) - float(sigma))).half()
    _, c_v, c_x, s_h, dt = sched.step_coefficients(sig.numpy(), step)
    ok = True
    for nhwc in (False, True):
        ua = u.permute(0, 1, 3, 4, 2).contiguous() if nhwc else u
        ca = c.permute(0, 1, 3, 4, 2).contiguous() if nhwc else c
        out = native.euler_vpred_step(torch.empty_like(lat), lat, ua, v_cond=ca if cfg else None,
                                      gs=gs.reshape(-1).contiguous() if cfg else None, v_nhwc=nhwc, c_v=c_v, c_x=c_x,
                                      sigma=s_h, dt=dt)
        ok = ok and torch.equal(out, ref)
        err = (out.float() - ref.float()).abs().max().item()
    torch.cuda.synchronize()
    return dict(max_err=err, tol=0.0, ok=bool(ok))


def handoff_flags(cfg=True):
    """The flag-signalled handoff on ONE device (set, then wait: a wait that had to block on another launch of the same
    GPU is exactly what must not be built): the Euler kernel with a handoff writes `out` and raises the flag from its
    last block; svdpp_flag_wait then returns at once, re-arms the flag, and the completion counter is back at zero."""
    B, C, Fr, H, W = 1, 4, 5, 24, 40
    lat = (_rand(B, C, Fr, H, W, seed=1).float() * 50).half()
    v = _rand(B, Fr, H, W, C, seed=2)
    vc = _rand(B, Fr, H, W, C, seed=3) if cfg else None
    gs = torch.linspace(1.0, 3.0, Fr, device=DEV).half() if cfg else None
    kw = dict(v_cond=vc, gs=gs, v_nhwc=True, c_v=-0.99, c_x=1.5, sigma=0.7, dt=-0.2)
    want = native.euler_vpred_step(torch.empty_like(lat), lat, v, **kw)
    flags = torch.zeros(4, dtype=torch.int32, device=DEV)
    done = torch.zeros(1, dtype=torch.int32, device=DEV)
    ok = True
    for rep in range(3):      # the counter re-arms itself: three hand-overs through the same flag
        out = torch.full_like(lat, float("nan"))
        native.euler_vpred_step(out, lat, v, handoff=(done.data_ptr(), flags[1:].data_ptr(), 1), **kw)
        native.flag_wait(flags[1:].data_ptr(), 1, reset_to=0, timeout_s=5)
        native.flag_set(flags[2:].data_ptr(), 7 + rep)
        torch.cuda.synchronize()
        ok = ok and torch.equal(out, want) and flags.tolist() == [0, 0, 7 + rep, 0] and int(done.item()) == 0
    return dict(max_err=0.0 if ok else 1.0, tol=0.0, ok=bool(ok))


def softmax_rows(rows=70, n=384, scale=0.3):
    x = (_rand(rows, n + 8, seed=1).float() * 6).half()
    got = x.clone()
    native.softmax_rows(got[:, :n], scale)
    ref = torch.softmax(x[:, :n].float() * scale, dim=-1)
    torch.cuda.synchronize()
    r = _cmp(got[:, :n], ref, rel=2e-3, floor=1e-5)
    r["ok"] = bool(r["ok"] and torch.equal(got[:, n:], x[:, n:]))          # columns beyond n (the row pitch) untouched
    return r


def transpose(R=300, C=520):
    x = _rand(R, C + 8, seed=1)[:, :C]
    out = torch.full((C, R + 8), float("nan"), device=DEV, dtype=torch.float16)
    native.transpose(out[:, :R], x)
    torch.cuda.synchronize()
    ok = torch.equal(out[:, :R], x.t()) and bool(torch.isnan(out[:, R:]).all())
    return dict(max_err=0.0 if ok else 1.0, tol=0.0, ok=bool(ok))


def time_conv_out(B=2, Fr=5, H=6, W=10, fp32=True):
    x = _rand(B * Fr * H * W, 4, seed=1)
    w, b = _rand(3, 3, 3, seed=2), _rand(3, seed=3)
    out = torch.full((B * Fr, 3, H, W), float("nan"), device=DEV, dtype=torch.float32 if fp32 else torch.float16)
    native.time_conv_out(out, x, w, b, B=B, F=Fr, HW=H * W)
    x5 = x[:, :3].float().reshape(B, Fr, H, W, 3).permute(0, 4, 1, 2, 3)                 # [B, 3, F, H, W]
    ref = F.conv3d(x5, w.float()[:, :, :, None, None], b.float(), padding=(1, 0, 0))
    ref = ref.permute(0, 2, 1, 3, 4).reshape(B * Fr, 3, H, W)
    torch.cuda.synchronize()
    return _cmp(out, ref, rel=2e-3, floor=2e-3)


BAYER4 = [[0, 8, 2, 10], [12, 4, 14, 6], [3, 11, 1, 9], [15, 7, 13, 5]]


def cube_index_reference(rgb_u8, dither=True):
    """numpy restatement of the fixed-palette quantiser of svdpp_frames_to_bytes (test oracle): rgb uint8 [F, H, W, 3] ->
    index uint8 [F, H, W] = r * 42 + g * 6 + b with level = min(L - 1, floor(v * (L - 1) / 255 + threshold)) in fp32."""
    import numpy as np
    Fr, H, W, _ = rgb_u8.shape
    thr = ((np.array(BAYER4, dtype=np.float32) + np.float32(0.5)) / np.float32(16)) if dither else np.full((4, 4), 0.5, np.float32)
    t = thr[np.arange(H)[:, None] & 3, np.arange(W)[None, :] & 3][None]                       # [1, H, W]
    lv = []
    for c, L in enumerate((6, 7, 6)):
        v = rgb_u8[..., c].astype(np.float32) * (np.float32(L - 1) / np.float32(255)) + t
        lv.append(np.minimum(v.astype(np.int32), L - 1))
    return (lv[0] * 42 + lv[1] * 6 + lv[2]).astype(np.uint8)


def frames_to_bytes(Fr=5, H=18, W=32, dtype=torch.float32, dither=True, permuted=True):
    """RGB bytes bit-equal to the reference script's torch expression (generate_video_demo.py:198-209), palette indices
    bit-equal to the numpy restatement; input as the permuted view decode_latents returns, values beyond [-1, 1]."""
    g = torch.Generator(device=DEV).manual_seed(5)
    base = (torch.rand(Fr, 3, H, W, device=DEV, generator=g) * 2.6 - 1.3).to(dtype)       # [F, 3, H, W] as the decoder writes it
    base[0, :, 0, :8] = torch.tensor([-1.0, 1.0, 0.0, -0.999, 0.999, 1.0 - 2.0 / 255, -2.0, 2.0], device=DEV, dtype=dtype)
    frames = base.permute(1, 0, 2, 3) if permuted else base.permute(1, 0, 2, 3).contiguous()   # [3, F, H, W]
    pad = 64
    rgb_buf = torch.full((Fr * H * W * 3 + 2 * pad,), 77, dtype=torch.uint8, device=DEV)
    rgb, idx = native.frames_to_bytes(frames, rgb=True, palette=True, dither=dither)
    only_idx = native.frames_to_bytes(frames, rgb=False, palette=True, dither=dither)[1]
    only_rgb = native.frames_to_bytes(frames, rgb=True, palette=False)[0]
    want = ((frames.float().permute(1, 2, 3, 0) + 1) / 2 * 255).clamp(0, 255).to(torch.uint8)      # the reference's expression
    torch.cuda.synchronize()
    ok_rgb = torch.equal(rgb, want) and torch.equal(only_rgb, want)
    want_idx = torch.from_numpy(cube_index_reference(want.cpu().numpy(), dither)).to(DEV)
    ok_idx = torch.equal(idx, want_idx) and torch.equal(only_idx, want_idx) and int(idx.max()) < 252
    # the palette entry of every index is within one cube step of the byte it stands for
    pal = torch.tensor(native.cube_palette(), device=DEV, dtype=torch.int32).reshape(256, 3)
    err = (pal[idx.long()] - want.int()).abs().amax().item()
    del rgb_buf
    return dict(max_err=float(err), tol=52.0, ok=bool(ok_rgb and ok_idx and err <= 52), rgb_equal=bool(ok_rgb), idx_equal=bool(ok_idx))


def _tiny_vae(seed=0, **over):
    from oracle.vae_torch import AutoencoderKLTemporalDecoder, tiny_vae_config
    from vdpp_b200.models.native_vae import NativeVAE
    torch.manual_seed(seed)
    cfg = tiny_vae_config(**over)
    oracle = AutoencoderKLTemporalDecoder(**cfg).to(DEV).eval()
    with torch.no_grad():                       # non-trivial norms / blends so that every path shows up in the output
        for n_, p_ in oracle.named_parameters():
            if n_.endswith("mix_factor"):
                p_.fill_(0.3)
            elif "norm" in n_ and n_.endswith("weight"):
                p_.add_(0.2 * torch.randn_like(p_))
            elif "norm" in n_ and n_.endswith("bias"):
                p_.add_(0.1 * torch.randn_like(p_))
    oracle = oracle.half()                      # fp16-representable weights on both sides
    nat = NativeVAE(oracle.state_dict(), config=cfg, device=DEV)
    return oracle, nat


def vae_decode(B=1, Fr=3, h=8, w=16, **over):
    """NativeVAE.decode vs the torch restatement of AutoencoderKLTemporalDecoder in fp32 on the same (fp16) weights."""
    oracle, nat = _tiny_vae(**over)
    z = (_rand(B * Fr, 4, h, w, seed=5).float() * 3).half()
    got = nat.decode(z, num_frames=Fr, out_dtype=torch.float32).sample
    with torch.no_grad():
        ref32 = oracle.float().decode(z.float(), num_frames=Fr).sample
        ref16 = oracle.half().decode(z, num_frames=Fr).sample
    torch.cuda.synchronize()
    r = _cmp(got, ref32, rel=2e-2, floor=2e-3)
    floor16 = (ref16.float() - ref32).abs().max().item()
    r["lib_fp16_vs_fp32"] = floor16
    r["ok"] = bool(r["ok"] and r["max_err"] <= max(4 * floor16, 5e-3) and tuple(got.shape) == tuple(ref32.shape))
    return r


def vae_encode(N=2, H=64, W=128, **over):
    oracle, nat = _tiny_vae(**over)
    x = (_rand(N, 3, H, W, seed=6).float().clamp(-1, 1)).half()
    got = nat.encode(x).latent_dist.mode()
    with torch.no_grad():
        ref32 = oracle.float().encode(x.float()).latent_dist.mode()
        ref16 = oracle.half().encode(x).latent_dist.mode()
    torch.cuda.synchronize()
    r = _cmp(got, ref32, rel=2e-2, floor=2e-3)
    floor16 = (ref16.float() - ref32).abs().max().item()
    r["lib_fp16_vs_fp32"] = floor16
    r["ok"] = bool(r["ok"] and r["max_err"] <= max(4 * floor16, 5e-3) and tuple(got.shape) == tuple(ref32.shape))
    return r


def attn_small(n_img=2, S=257, S_pad=384, heads=3, hd=80, stride=128, ostride=128):
    """svdpp_attn_small_f16 vs torch fp32 SDPA: padded head layout, padding tokens / padding columns written as zeros."""
    width = 3 * heads * stride
    qkv = _rand(n_img * S_pad, width, seed=5)
    guard = _Guarded(n_img * S_pad, heads * ostride)
    out = guard.out
    out.fill_(float("nan"))              # every element of the output must be written
    native.attn_small(out, qkv, n_img=n_img, S=S, S_pad=S_pad, heads=heads, head_dim=hd, q_off=0, k_off=heads * stride,
                      v_off=2 * heads * stride, head_stride=stride, out_head_stride=ostride, scale=hd ** -0.5)
    t = qkv.reshape(n_img, S_pad, 3, heads, stride)[:, :S, :, :, :hd].float()      # [n, S, 3, h, hd]
    q, k, v = [t[:, :, i].transpose(1, 2) for i in range(3)]                        # [n, h, S, hd]
    ref = torch.zeros(n_img, S_pad, heads, ostride, device=DEV)
    ref[:, :S, :, :hd] = F.scaled_dot_product_attention(q, k, v).transpose(1, 2)
    torch.cuda.synchronize()
    return _with_guard(_cmp(out, ref.reshape(n_img * S_pad, heads * ostride), rel=2e-3), guard)


def clip_vision(B=2, **over):
    """NativeCLIPVision vs the REAL transformers CLIPVisionModelWithProjection on the same weights (fp32 and fp16)."""
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    from vdpp_b200.models.native_clip import NativeCLIPVision
    cfg = dict(hidden_size=192, intermediate_size=384, num_hidden_layers=2, num_attention_heads=4, image_size=56,
               patch_size=14, projection_dim=64, hidden_act="gelu")
    cfg.update(over)
    torch.manual_seed(0)
    attn, use_graph = cfg.pop("attn", "fused"), cfg.pop("use_graph", False)
    lib = CLIPVisionModelWithProjection(CLIPVisionConfig(**cfg)).to(DEV).half().eval()
    nat = NativeCLIPVision(lib.state_dict(), config=cfg, device=DEV, attn=attn, use_graph=use_graph)
    px = _rand(B, 3, cfg["image_size"], cfg["image_size"], seed=3)
    if use_graph:      # capture on other pixels, then replay on the checked ones
        nat(_rand(B, 3, cfg["image_size"], cfg["image_size"], seed=4))
    got = nat(px).image_embeds
    with torch.no_grad():
        ref16 = lib(pixel_values=px).image_embeds
        ref32 = lib.float()(pixel_values=px.float()).image_embeds
    torch.cuda.synchronize()
    r = _cmp(got, ref32, rel=1e-2, floor=2e-3)
    floor16 = (ref16.float() - ref32).abs().max().item()
    r["lib_fp16_vs_fp32"] = floor16
    r["ok"] = bool(r["ok"] and r["max_err"] <= max(4 * floor16, 3e-3) and tuple(got.shape) == tuple(ref32.shape))
    return r


def dummy_unet(C=4, Ch=16, step=7):
    from vdpp_b200.models import DummyUNet
    torch.manual_seed(0)
    m = DummyUNet(channels=C, hidden_channels=Ch).to(DEV)
    x = torch.randn(1, C, 6, 16, 16, device=DEV)
    m.native_cuda = False                     # the torch arithmetic of the reference's module
    with torch.no_grad():
        ref = m(x, step)
        m.native_cuda = True                  # ... and the module's own dispatch to the native kernel
        via_module = m(x, step)
    assert (via_module - ref).abs().max().item() <= 1e-4 * ref.abs().max().item() + 1e-4
    hid = torch.empty(1, Ch, 6, 16, 16, device=DEV)
    out = native.dummy_unet_step(torch.empty_like(x), x, m.net[0].weight.contiguous(), m.net[0].bias,
                                 m.net[2].weight.contiguous(), m.net[2].bias, m.norm.weight, m.norm.bias, m.norm.eps,
                                 math.tanh(step / 10.0), hid)
    torch.cuda.synchronize()
    return _cmp(out, ref, rel=1e-4, floor=1e-4)


def _tuned_later(fn_name, kw, **switches):
    """A check that runs globals()[fn_name](**kw) with tuning switches set (restored afterwards)."""
    def run():
        old = {k: native.set_tuning(k, v) for k, v in switches.items()}
        try:
            return globals()[fn_name](**kw)
        finally:
            torch.cuda.synchronize()
            for k, v in old.items():
                native.set_tuning(k, v)
    return run


ALL_CHECKS = {
    "movers": lambda: movers(),
    "euler_cfg": lambda: euler(True),
    "euler_nocfg": lambda: euler(False),
    "sinusoid": lambda: sinusoid(),
    "linear_small": lambda: linear_small(),
    "layernorm": lambda: layernorm(),
    "layernorm_1280": lambda: layernorm(M=33, C=1280, add=False),
    "layernorm_640": lambda: layernorm(M=37, C=640, add=True),
    "layernorm_generic": lambda: layernorm(M=19, C=128, add=True),
    "softmax_rows": lambda: softmax_rows(),
    "softmax_rows_9216": lambda: softmax_rows(rows=33, n=9216, scale=512 ** -0.5),
    # whole short-sequence attention in one launch (the CLIP image encoder's 257 tokens x 16 heads of width 80)
    "attn_small_clip": lambda: attn_small(),
    "attn_small_tiny": lambda: attn_small(n_img=1, S=17, S_pad=128, heads=4, hd=48, stride=128, ostride=128),
    "attn_small_dense": lambda: attn_small(n_img=3, S=64, S_pad=64, heads=2, hd=64, stride=64, ostride=64),
    "attn_small_hd128": lambda: attn_small(n_img=1, S=288, S_pad=288, heads=2, hd=128, stride=128, ostride=128),
    "attn_small_long": lambda: attn_small(n_img=2, S=400, S_pad=512, heads=2, hd=40, stride=64, ostride=48),
    "transpose": lambda: transpose(),
    "transpose_9216x512": lambda: transpose(R=9216, C=512),
    "frames_to_bytes_fp32": lambda: frames_to_bytes(),
    "frames_to_bytes_fp16_contiguous": lambda: frames_to_bytes(Fr=1, H=7, W=12, dtype=torch.float16, permuted=False),
    "frames_to_bytes_nodither_large": lambda: frames_to_bytes(Fr=3, H=576, W=1024, dither=False),
    "time_conv_out_fp32": lambda: time_conv_out(fp32=True),
    "time_conv_out_fp16": lambda: time_conv_out(B=1, Fr=1, fp32=False),
    "handoff_flags_cfg": lambda: handoff_flags(True),
    "handoff_flags_nocfg": lambda: handoff_flags(False),
    "groupnorm": lambda: groupnorm(),
    # variance ~ 4e-6: eps = 1e-6 and 1e-5 give outputs 1.7x apart, so the scalar must reach the kernel unchanged
    "groupnorm_small_var_eps_1e6": lambda: groupnorm(eps=1e-6, in_scale=1e-3),
    "groupnorm_small_var_eps_1e5": lambda: groupnorm(eps=1e-5, in_scale=1e-3),
    "groupnorm_cat": lambda: groupnorm(n_img=2, HW=144, C1=128, C2=64),
    "groupnorm_temporal": lambda: groupnorm(n_img=6, HW=50, C1=320, fps=3, eps=1e-6, silu=False),
    "attn_temporal": lambda: attn_temporal(),
    "attn_temporal_25": lambda: attn_temporal(B=1, Fr=25, HW=9, heads=5),
    "attn_temporal_25_many": lambda: attn_temporal(B=2, Fr=25, HW=2304, heads=10),
    "attn_temporal_14": lambda: attn_temporal(B=1, Fr=14, HW=577, heads=5),
    "attn_temporal_1": lambda: attn_temporal(B=2, Fr=1, HW=33, heads=1),
    "attn_temporal_8": lambda: attn_temporal(B=1, Fr=8, HW=100, heads=3),
    "attn_temporal_17": lambda: attn_temporal(B=1, Fr=17, HW=100, heads=2),
    "attn_temporal_32": lambda: attn_temporal(B=1, Fr=32, HW=64, heads=20),
    "linear_small_grouped": lambda: linear_small_grouped(),
    "groupnorm_1280": lambda: groupnorm(n_img=3, HW=144, C1=1280, eps=1e-6),
    "groupnorm_cat_1920": lambda: groupnorm(n_img=2, HW=576, C1=1280, C2=640),
    "groupnorm_temporal_big": lambda: groupnorm(n_img=10, HW=2304, C1=640, fps=5),
    "dummy_unet": lambda: dummy_unet(),
    "simt_gemm_linear": lambda: gemm_linear(impl=1),
    "simt_gemm_split": lambda: gemm_linear(impl=1, K=384, split=True),
    "simt_gemm_geglu": lambda: gemm_geglu(impl=1),
    "simt_conv3x3": lambda: conv3x3(impl=1),
    "simt_conv_temporal": lambda: conv_temporal(impl=1),
    "simt_attn_spatial": lambda: attn_spatial(impl=1),
    "tc_gemm_plain": lambda: gemm_linear(M=256, N=160, K=64, impl=0, epilogue="bias"),
    "tc_gemm_linear": lambda: gemm_linear(impl=0),
    "tc_gemm_big": lambda: gemm_linear(M=4000, N=640, K=1280, impl=0),
    "tc_gemm_split": lambda: gemm_linear(impl=0, K=384, split=True),
    "tc_gemm_geglu": lambda: gemm_geglu(impl=0),
    "tc_gemm_geglu_320": lambda: gemm_geglu(M=1000, C=320, impl=0),
    "tc_conv3x3_w32": lambda: conv3x3(impl=0),
    "tc_conv3x3_w128": lambda: conv3x3(B=1, Fr=2, H=3, W=128, C=128, Cout=160, impl=0),
    "tc_conv3x3_w16": lambda: conv3x3(B=1, Fr=3, H=9, W=16, C=64, Cout=64, impl=0),
    "tc_conv_temporal": lambda: conv_temporal(impl=0),
    "simt_conv_up2x": lambda: conv_up2x(impl=1),
    "tc_conv_up2x_w32": lambda: conv_up2x(impl=0),
    "tc_conv_up2x_w16": lambda: conv_up2x(B=1, Fr=3, H=9, W=16, C=64, Cout=64, impl=0),
    "pair256_conv_up2x": lambda: conv_up2x(B=2, Fr=2, H=8, W=32, C=128, Cout=256, impl=3),
    "pair320_conv_up2x": lambda: conv_up2x(B=2, Fr=2, H=8, W=64, C=128, Cout=320, impl=6),
    "simt_conv3x3_stride2": lambda: conv3x3_stride2(impl=1),
    "tc_conv3x3_stride2_w32": lambda: conv3x3_stride2(impl=0),
    "tc_conv3x3_stride2_w16_odd": lambda: conv3x3_stride2(B=1, Fr=3, H=9, W=32, C=64, Cout=64, impl=0),
    "tc_conv3x3_stride2_w128": lambda: conv3x3_stride2(B=1, Fr=1, H=6, W=256, C=64, Cout=160, impl=0),
    "pair256_conv3x3_stride2": lambda: conv3x3_stride2(B=2, Fr=2, H=16, W=64, C=128, Cout=256, impl=3),
    "pair320_conv3x3_stride2": lambda: conv3x3_stride2(B=2, Fr=2, H=16, W=64, C=128, Cout=320, impl=6),
    "pair_gemm_plain": lambda: gemm_linear(M=256, N=160, K=64, impl=2, epilogue="bias"),
    "pair_gemm_linear": lambda: gemm_linear(impl=2),
    "pair_gemm_big": lambda: gemm_linear(M=4000, N=640, K=1280, impl=2),
    "pair_gemm_split": lambda: gemm_linear(impl=2, K=384, split=True),
    "pair_gemm_geglu_320": lambda: gemm_geglu(M=1000, C=320, impl=2),
    "pair_conv3x3_w32": lambda: conv3x3(impl=2),
    "pair_conv3x3_w128": lambda: conv3x3(B=1, Fr=2, H=3, W=128, C=128, Cout=160, impl=2),
    "pair_conv_temporal": lambda: conv_temporal(impl=2),
    "pair256_gemm_plain": lambda: gemm_linear(M=512, N=256, K=64, impl=3, epilogue="bias"),
    "pair256_gemm_linear": lambda: gemm_linear(M=300, N=1280, K=320, impl=3),
    "pair256_gemm_big": lambda: gemm_linear(M=4000, N=1280, K=1280, impl=3),
    "pair256_gemm_split": lambda: gemm_linear(M=700, N=512, impl=3, K=384, split=True),
    "pair256_gemm_geglu_320": lambda: gemm_geglu(M=1000, C=320, impl=3),
    "pair256_conv3x3_w32": lambda: conv3x3(Cout=256, impl=3),
    "pair256_conv_temporal": lambda: conv_temporal(B=2, Fr=5, H=4, W=32, C=256, impl=3),
    "bn128_gemm_plain": lambda: gemm_linear(M=256, N=128, K=64, impl=4, epilogue="bias"),
    "bn128_gemm_linear": lambda: gemm_linear(M=300, N=1280, K=320, impl=4),
    "bn128_gemm_big": lambda: gemm_linear(M=3600, N=1280, K=1280, impl=4),
    "bn128_gemm_split": lambda: gemm_linear(M=700, N=512, impl=4, K=384, split=True),
    "bn128_conv3x3_w16": lambda: conv3x3(B=1, Fr=3, H=9, W=16, C=64, Cout=128, impl=4),
    "bn128_conv_temporal": lambda: conv_temporal(B=2, Fr=5, H=4, W=32, C=128, impl=4),
    "pair128_gemm_plain": lambda: gemm_linear(M=256, N=128, K=64, impl=7, epilogue="bias"),
    "pair128_gemm_linear": lambda: gemm_linear(M=300, N=1280, K=320, impl=7),
    "pair128_gemm_big": lambda: gemm_linear(M=3600, N=1280, K=1280, impl=7),
    "pair128_gemm_mtail_odd": lambda: gemm_linear(M=385, N=256, K=128, impl=7),
    "pair128_gemm_many_tiles": lambda: gemm_linear(M=70000, N=128, K=384, impl=7),
    "pair128_gemm_split": lambda: gemm_linear(M=700, N=512, impl=7, K=384, split=True),
    "pair128_conv3x3_w16": lambda: conv3x3(B=1, Fr=3, H=9, W=16, C=64, Cout=128, impl=7),
    "pair128_conv3x3_w64": lambda: conv3x3(B=1, Fr=2, H=24, W=64, C=128, Cout=128, impl=7),
    "pair128_conv_temporal": lambda: conv_temporal(B=2, Fr=5, H=4, W=32, C=128, impl=7),
    "pair256_gemm_nstore_partial": lambda: gemm_linear(M=700, N=960, K=320, impl=3, epilogue="full"),
    "pair256_gemm_nstore_partial_bias": lambda: gemm_linear(M=300, N=1920, K=640, impl=3, epilogue="bias"),
    "pair320_gemm_plain": lambda: gemm_linear(M=512, N=320, K=64, impl=6, epilogue="bias"),
    "pair320_gemm_linear": lambda: gemm_linear(M=300, N=320, K=320, impl=6),
    "pair320_gemm_big": lambda: gemm_linear(M=4000, N=640, K=1280, impl=6),
    "pair320_gemm_split": lambda: gemm_linear(M=700, N=320, impl=6, K=384, split=True),
    "pair320_conv3x3_w32": lambda: conv3x3(Cout=320, impl=6),
    "pair320_conv3x3_w128": lambda: conv3x3(B=1, Fr=2, H=3, W=128, C=128, Cout=320, impl=6),
    "pair320_conv_temporal": lambda: conv_temporal(B=2, Fr=5, H=4, W=32, C=320, impl=6),
    "tc_attn_spatial_256": lambda: attn_spatial(n_img=1, S=256, heads=1, impl=0),
    "tc_attn_spatial_tail": lambda: attn_spatial(n_img=2, S=320, heads=2, impl=0),
    "tc_attn_spatial_144": lambda: attn_spatial(n_img=3, S=144, heads=2, impl=0),
    "tc_attn_spatial_2304": lambda: attn_spatial(n_img=2, S=2304, heads=5, impl=0),
    "tc_attn_spatial_rescale": lambda: attn_spatial(n_img=2, S=1000, heads=2, impl=0, growing=True),
    "tc2_attn_spatial_256": lambda: attn_spatial(n_img=1, S=256, heads=1, impl=2),
    "tc2_attn_spatial_tail": lambda: attn_spatial(n_img=2, S=320, heads=2, impl=2),
    "tc2_attn_spatial_144": lambda: attn_spatial(n_img=3, S=144, heads=2, impl=2),
    "tc2_attn_spatial_576": lambda: attn_spatial(n_img=2, S=576, heads=3, impl=2),
    "tc2_attn_spatial_2304": lambda: attn_spatial(n_img=2, S=2304, heads=5, impl=2),
    "tc2_attn_spatial_rescale": lambda: attn_spatial(n_img=2, S=1000, heads=2, impl=2, growing=True),
    "tc3_attn_spatial_256": lambda: attn_spatial(n_img=1, S=256, heads=1, impl=3),
    "tc3_attn_spatial_tail": lambda: attn_spatial(n_img=2, S=320, heads=2, impl=3),
    "tc3_attn_spatial_144": lambda: attn_spatial(n_img=3, S=144, heads=2, impl=3),
    "tc3_attn_spatial_576": lambda: attn_spatial(n_img=2, S=576, heads=3, impl=3),
    "tc3_attn_spatial_2304": lambda: attn_spatial(n_img=2, S=2304, heads=5, impl=3),
    "tc3_attn_spatial_rescale": lambda: attn_spatial(n_img=2, S=1000, heads=2, impl=3, growing=True),
    "simt_attn_spatial_rescale": lambda: attn_spatial(n_img=1, S=600, heads=1, impl=1, growing=True),
    # quarter-pipelined softmax (impl 4; 5 / 6: every 8th / 4th group of exponentials as an FMA-pipe polynomial)
    **{f"tc{i}_attn_spatial_{n}": (lambda i=i, kw=kw: attn_spatial(impl=i, **kw))
       for i in (4, 5, 6)
       for n, kw in (("256", dict(n_img=1, S=256, heads=1)), ("tail", dict(n_img=2, S=320, heads=2)),
                     ("144", dict(n_img=3, S=144, heads=2)), ("576", dict(n_img=2, S=576, heads=3)),
                     ("2304", dict(n_img=2, S=2304, heads=5)), ("rescale", dict(n_img=2, S=1000, heads=2, growing=True)),
                     ("rescale_late", dict(n_img=1, S=1200, heads=1, growing="late")))},
    # fused GEGLU feed-forward (ff_fused.cu): the gated intermediate stays in tensor memory
    "ff_fused_320_res": lambda: ff_fused(),
    "ff_fused_320_plain": lambda: ff_fused(M=128 * 3, epilogue="none"),
    "ff_fused_320_blend": lambda: ff_fused(M=40000, epilogue="blend"),      # ~2 tiles per CTA
    "ff_fused_320_rowvec": lambda: ff_fused(M=777, epilogue="rowvec"),
    "ff_fused_64": lambda: ff_fused(M=300, C=64),
    "ff_fused_256_many": lambda: ff_fused(M=128 * 148 * 3 + 5, C=256),      # 3+ tiles per CTA, M tail
    "ff_fused_320_odd_tiles": lambda: ff_fused(M=128 * 7 + 3, epilogue="blend"),   # odd tile count: one CTA of the last pair idles
    **{f"ff_fused_{n}_mc": _tuned_later("ff_fused", kw, ff_pair=1)       # TMA-multicast pairs with independent MMAs
       for n, kw in (("320_res", dict()), ("320_blend", dict(M=40000, epilogue="blend")), ("320_odd_tiles", dict(M=128 * 7 + 3)))},
    **{f"ff_fused_{n}_two": _tuned_later("ff_fused", kw, ff_pair=2)      # cta_group::2 pairs (the default): the leader issues M = 256 MMAs
       for n, kw in (("320_res", dict()), ("320_blend", dict(M=40000, epilogue="blend")), ("64", dict(M=300, C=64)),
                     ("320_odd_tiles", dict(M=128 * 7 + 3, epilogue="rowvec")), ("256_many", dict(M=128 * 148 * 3 + 5, C=256)))},
    **{f"ff_fused_{n}_nopair": _tuned_later("ff_fused", kw, ff_pair=0)
       for n, kw in (("320_res", dict()), ("320_blend", dict(M=40000, epilogue="blend")), ("64", dict(M=300, C=64)))},
    # ping-pong kernel (impl 7, fmha3_tc.cu): the two tiles' exponential phases alternate through named barriers
    **{f"tc7_attn_spatial_{n}": (lambda kw=kw: attn_spatial(impl=7, **kw))
       for n, kw in (("64", dict(n_img=2, S=64, heads=1)), ("256", dict(n_img=1, S=256, heads=1)),
                     ("tail", dict(n_img=2, S=320, heads=2)), ("144", dict(n_img=3, S=144, heads=2)),
                     ("576", dict(n_img=2, S=576, heads=3)), ("2304", dict(n_img=2, S=2304, heads=5)),
                     ("rescale", dict(n_img=2, S=1000, heads=2, growing=True)),
                     ("rescale_late", dict(n_img=1, S=1200, heads=1, growing="late")))},
    # two threads per row, the two tiles' half-row warps take turns in pairs (impl 8, fmha4_tc.cu)
    **{f"tc8_attn_spatial_{n}": (lambda kw=kw: attn_spatial(impl=8, **kw))
       for n, kw in (("64", dict(n_img=2, S=64, heads=1)), ("256", dict(n_img=1, S=256, heads=1)),
                     ("tail", dict(n_img=2, S=320, heads=2)), ("144", dict(n_img=3, S=144, heads=2)),
                     ("200", dict(n_img=2, S=200, heads=1)),
                     ("576", dict(n_img=2, S=576, heads=3)), ("2304", dict(n_img=2, S=2304, heads=5)),
                     ("rescale", dict(n_img=2, S=1000, heads=2, growing=True)),
                     ("rescale_late", dict(n_img=1, S=1200, heads=1, growing="late")))},
    **{f"tc8_attn_spatial_handover{h}": _tuned_later("attn_spatial", dict(impl=8, n_img=2, S=1000, heads=2, growing=True),
                                                     fmha_handover_split=h) for h in (0, 3)},
    **{f"tc7_attn_spatial_handover{h}": _tuned_later("attn_spatial", dict(impl=7, n_img=2, S=1000, heads=2, growing=True), fmha_handover=h)
       for h in (0, 7)},
    # a share of the exponentials on the FMA pipe (degree-3 polynomial, "fmha_poly" of every 16): every instantiation,
    # plain / partial last block with masked keys / the lazy-rescale path
    **{f"tc7_attn_spatial_poly{n}_{tag}": _tuned_later("attn_spatial", dict(impl=7, **kw), fmha_poly=n)
       for n in (0, 2, 3, 4, 5, 6, 8)
       for tag, kw in (("2304", dict(n_img=2, S=2304, heads=2)), ("tail", dict(n_img=2, S=1000, heads=2, growing=True)),
                       ("144", dict(n_img=3, S=144, heads=2)))},
}


def _tuned(fn, **switches):
    """Run a check with kernel tuning switches (svdpp_set_tuning) set, then restore them."""
    def run():
        old = {k: native.set_tuning(k, v) for k, v in switches.items()}
        try:
            return fn()
        finally:
            torch.cuda.synchronize()
            for k, v in old.items():
                native.set_tuning(k, v)
    return run


def pdl_chain(n=6):
    """A chain of dependent launches (GEMM -> LayerNorm -> GEMM -> GroupNorm -> ...) must give bit-identical
    results with and without programmatic dependent launch: every kernel waits for its predecessor's writes."""
    M, C = 2304, 320
    x0 = _rand(M, C, seed=11)
    w = _pad_n(_rand(C, C, scale=C ** -0.5, seed=12))
    bias = _rand(C, seed=13)
    gam, bet = _rand(C, seed=14), _rand(C, seed=15)
    ws = torch.zeros((native.groupnorm_workspace_bytes(4, M // 4) + 3) // 4, dtype=torch.float32, device=DEV)

    def run():
        x = x0
        for i in range(n):
            y = torch.empty_like(x)
            native.gemm(y, x, w, bias=bias, r1=x, beta1=0.5, alpha=0.5, n_store=C, impl=(0, 2, 6)[i % 3])
            z = torch.empty_like(x)
            native.layernorm(z, y, gam, bet)
            x = torch.empty_like(x)
            native.groupnorm_silu(x, z, gam, bet, n_img=4, HW=M // 4, eps=1e-5, silu=True, workspace=ws)
        torch.cuda.synchronize()
        return x

    old = native.set_tuning("pdl", 0)
    try:
        a = run()
        native.set_tuning("pdl", 1)
        b = run()
        c = run()
    finally:
        native.set_tuning("pdl", old)
    same = bool(torch.equal(a, b) and torch.equal(b, c))
    return dict(max_err=(a.float() - b.float()).abs().max().item(), tol=0.0, ok=bool(same and torch.isfinite(a).all()))


_LEGACY = ("tc_gemm_plain", "tc_gemm_linear", "tc_gemm_big", "tc_gemm_split", "tc_conv3x3_w32", "tc_conv_temporal",
           "pair_gemm_linear", "pair_gemm_big", "pair256_gemm_linear", "pair256_gemm_big", "pair256_gemm_geglu_320",
           "pair256_gemm_nstore_partial", "pair256_gemm_nstore_partial_bias", "bn128_gemm_linear", "pair320_gemm_linear",
           "pair320_gemm_big", "pair320_conv3x3_w32", "pair256_conv_temporal")
for _n in _LEGACY:                       # the per-thread copy-out path (tma_store = 0) stays covered
    ALL_CHECKS["copyout_" + _n] = _tuned(ALL_CHECKS[_n], tma_store=0)
for _n in ("tc_gemm_linear", "pair256_gemm_big", "pair320_gemm_big", "pair256_gemm_geglu_320", "tc2_attn_spatial_2304",
           "tc_attn_spatial_tail", "attn_temporal_25_many", "groupnorm_cat_1920", "layernorm_640", "euler_cfg"):
    ALL_CHECKS["pdl_" + _n] = _tuned(ALL_CHECKS[_n], pdl=1)
ALL_CHECKS["pdl_chain"] = pdl_chain
ALL_CHECKS["pair256_gemm_mtail_odd"] = lambda: gemm_linear(M=385, N=512, K=128, impl=3)   # odd tile count: one CTA of the last pair idles
ALL_CHECKS["pair320_gemm_mtail_odd"] = lambda: gemm_linear(M=385, N=640, K=128, impl=6)
for _n in ("tc_gemm_linear", "tc_gemm_big", "tc_gemm_split", "pair_gemm_big", "pair256_gemm_linear", "pair256_gemm_big",
           "pair256_gemm_nstore_partial", "pair320_gemm_linear", "pair320_gemm_big", "bn128_gemm_big",
           "pair256_gemm_mtail_odd", "pair320_gemm_mtail_odd"):
    ALL_CHECKS["r1ldg_" + _n] = _tuned(ALL_CHECKS[_n], tma_r1=0)       # residual tile through per-thread loads (TMA loads are the default)
for _n in ("tc_gemm_big", "pair_gemm_big", "pair256_gemm_big", "pair320_gemm_big", "bn128_gemm_big", "pair128_gemm_big",
           "pair320_conv3x3_w32", "pair256_conv_temporal", "tc_conv3x3_w32"):
    ALL_CHECKS["oneprod_" + _n] = _tuned(ALL_CHECKS[_n], two_prod=0)    # one TMA producer warp (two are the default for K >= 384)
ALL_CHECKS["tc_gemm_many_tiles"] = lambda: gemm_linear(M=40000, N=640, K=320, impl=0)          # ~8 tiles per CTA
ALL_CHECKS["pair256_gemm_many_tiles"] = lambda: gemm_linear(M=40000, N=1024, K=320, impl=3)
ALL_CHECKS["pair320_gemm_many_tiles"] = lambda: gemm_linear(M=70000, N=640, K=320, impl=6)
ALL_CHECKS["pair_gemm_many_tiles_bias"] = lambda: gemm_linear(M=40000, N=640, K=640, impl=2, epilogue="bias")
for _n in ("tc_gemm_plain", "tc_gemm_linear", "tc_gemm_split", "pair_gemm_plain", "pair_gemm_linear", "pair_gemm_split",
           "pair256_gemm_plain", "pair256_gemm_linear", "pair256_gemm_split", "pair256_gemm_nstore_partial",
           "pair256_gemm_nstore_partial_bias", "bn128_gemm_plain", "bn128_gemm_linear", "bn128_gemm_split",
           "pair320_gemm_plain", "pair320_gemm_linear", "pair320_gemm_split", "pair256_gemm_mtail_odd",
           "pair320_gemm_mtail_odd", "tc_conv3x3_w32", "tc_conv3x3_w16", "tc_conv_temporal", "pair_conv3x3_w32",
           "pair256_conv3x3_w32", "pair320_conv3x3_w32", "bn128_conv3x3_w16", "tc_conv3x3_stride2_w32",
           "tc_gemm_many_tiles", "pair256_gemm_many_tiles", "pair320_gemm_many_tiles", "pair_gemm_many_tiles_bias"):
    ALL_CHECKS["dma_" + _n] = _tuned(ALL_CHECKS[_n], epi_dma=2)        # DMA-lane epilogue forced on every tile shape
for _n in ("pair256_gemm_linear", "pair256_gemm_nstore_partial", "pair256_gemm_many_tiles", "bn128_gemm_linear"):
    ALL_CHECKS["nodma_" + _n] = _tuned(ALL_CHECKS[_n], epi_dma=0)      # thread-0 epilogue where the DMA lane is the default
# split-K tail of the 256x320 pair tiles: 80 / 86 / 111 pair tiles on 74 pairs -> 6 / 12 / 37 tiles in the last wave
ALL_CHECKS["pair320_splitk_12way"] = lambda: gemm_linear(M=20480, N=320, K=3072, impl=6)
ALL_CHECKS["pair320_splitk_5way"] = lambda: gemm_linear(M=20480, N=320, K=1280, impl=6)
ALL_CHECKS["pair320_splitk_6way_bias"] = lambda: gemm_linear(M=11008, N=640, K=1920, impl=6, epilogue="bias")
ALL_CHECKS["pair320_splitk_2way_mtail"] = lambda: gemm_linear(M=28400, N=320, K=640, impl=6)
ALL_CHECKS["pair320_splitk_conv"] = lambda: conv3x3(B=1, Fr=10, H=64, W=32, C=128, Cout=320, impl=6)
ALL_CHECKS["pair320_splitk_conv_temporal"] = lambda: conv_temporal(B=2, Fr=5, H=64, W=32, C=320, impl=6)
for _n in ("pair320_splitk_12way", "pair320_splitk_5way", "pair320_splitk_6way_bias", "pair320_splitk_2way_mtail",
           "pair320_splitk_conv", "pair320_splitk_conv_temporal"):
    ALL_CHECKS[_n] = _tuned(ALL_CHECKS[_n], splitk_min_total_kb=1)      # small K here: lift the K threshold of the default rule
for _n in ("pair320_splitk_12way", "pair320_splitk_conv"):
    ALL_CHECKS["nosplit_" + _n] = _tuned(ALL_CHECKS[_n], splitk=0)
ALL_CHECKS["tc_gemm_geglu_tail"] = lambda: gemm_geglu(M=1000, C=320, impl=3)
# traversal direction ("reverse": rows walked from the end - the L2-friendly order after a forward producer); results must not
# depend on it.  GroupNorm: 1 = statistics from the end / apply forward, 2 = the other way round
for _n in ("tc_gemm_linear", "tc_gemm_many_tiles", "pair256_gemm_many_tiles", "pair320_gemm_many_tiles", "pair256_gemm_mtail_odd",
           "pair320_gemm_mtail_odd", "pair256_gemm_geglu_320", "tc_conv3x3_w32", "pair320_conv3x3_w32", "tc_conv_temporal",
           "tc_conv3x3_stride2_w32", "tc_conv_up2x_w32", "pair320_splitk_12way", "pair320_splitk_conv", "bn128_gemm_linear",
           "layernorm", "layernorm_640", "layernorm_1280", "layernorm_generic", "groupnorm", "groupnorm_cat",
           "groupnorm_cat_1920"):
    if _n in ALL_CHECKS:
        ALL_CHECKS["rev_" + _n] = _tuned(ALL_CHECKS[_n], reverse=1)
for _n in ("groupnorm", "groupnorm_cat", "groupnorm_cat_1920"):
    if _n in ALL_CHECKS:
        ALL_CHECKS["rev2_" + _n] = _tuned(ALL_CHECKS[_n], reverse=2)


# ------------------------------------------------------------------------------------------ whole UNet
def _tiny_pair(cfg_over=None, gemm_impl=0, attn_impl=None, seed=0, orchestrator=None):
    from oracle.unet_torch import UNetSpatioTemporalConditionModel, tiny_config
    from vdpp_b200.models.native_unet import NativeUNet
    cfg = tiny_config(**(cfg_over or {}))
    torch.manual_seed(seed)
    oracle = UNetSpatioTemporalConditionModel(**cfg).to(DEV).half().eval()
    nat = NativeUNet(oracle.state_dict(), config=oracle.config, device=DEV, gemm_impl=gemm_impl, attn_impl=attn_impl,
                     orchestrator=orchestrator)
    return oracle, nat


def unet_tiny(gemm_impl=0, attn_impl=None, B=1, Fr=3, H=16, W=16, cfg_over=None, orchestrator=None):
    """NativeUNet vs the torch oracle (fp16 library kernels) and vs the oracle in fp32, same weights."""
    oracle, nat = _tiny_pair(cfg_over, gemm_impl, attn_impl, orchestrator=orchestrator)
    g = torch.Generator(device=DEV)
    g.manual_seed(1)
    sample = torch.randn(B, Fr, 8, H, W, device=DEV, generator=g).half()
    enc = torch.randn(B, 1, 1024, device=DEV, generator=g).half()
    ids = torch.tensor([[5.0, 127.0, 0.02]], device=DEV).half().repeat(B, 1)
    t = torch.tensor(1.6377)
    with torch.no_grad():
        ref16 = oracle(sample, t, enc, ids)[0]
        ref32 = oracle.float()(sample.float(), t, enc.float(), ids.float())[0]
        oracle.half()
    got = nat(sample, t, enc, ids)[0]
    torch.cuda.synchronize()
    r = _cmp(got, ref32, rel=2e-2, floor=2e-3)
    floor16 = (ref16.float() - ref32).abs().max().item()
    r["lib_fp16_vs_fp32"] = floor16
    r["native_vs_lib_fp16"] = (got.float() - ref16.float()).abs().max().item()
    # native must be about as close to the fp32 truth as the library fp16 path is
    r["ok"] = bool(r["ok"] and r["max_err"] <= max(4 * floor16, 5e-3))
    return r


def svd_steps(n_steps=4, total=25, cfg_scale=None, gemm_impl=0, attn_impl=None, B=1, Fr=3, H=16, W=16, graph=False,
              orchestrator=None):
    """StableVideoUNet (native) vs the oracle restatement of the reference wrapper, a few Euler steps."""
    from oracle.svd_step import OracleStep, dummy_conditioning
    from vdpp_b200.models import StableVideoUNet
    oracle, nat = _tiny_pair(None, gemm_impl, attn_impl, orchestrator=orchestrator)
    ts = StableVideoUNet._default_timestep_schedule(total)
    model = StableVideoUNet(unet=nat, timesteps=ts).to(DEV)
    model.use_cuda_graph = graph
    torch.manual_seed(5)
    cond = dummy_conditioning(B, Fr, H, W, torch.device(DEV), torch.float16, guidance_scale=cfg_scale)
    model.set_conditioning(cond.image_embeddings, cond.image_latents, guidance_scale=cfg_scale, num_frames=Fr)
    ostep = OracleStep(oracle, total)
    torch.manual_seed(42)
    x0 = torch.randn(B, 4, Fr, H, W, device=DEV).half() * model.init_noise_sigma
    a, b = x0.clone(), x0.clone()
    worst = 0.0
    rep = 2 if graph else 1
    for _ in range(rep):          # with graphs: first pass warms + captures, second replays
        a, b = x0.clone(), x0.clone()
        for s in range(n_steps):
            a = model(a, s)
            b = ostep(b, s, cond)
    torch.cuda.synchronize()
    scale = b.float().abs().max().item()
    # the latent is O(700) at these steps, where an fp16 ulp is 0.5: allow 3 ulps of the largest element after four
    # steps (observed: half an ulp) - a 1 % kernel error (8 at this scale) must not pass
    r = _cmp(a, b, rel=3.0 * 2.0 ** -10, floor=0.0)
    r["latent_absmax"] = scale
    return r


UNET_CHECKS = {
    "unet_tiny_simt": lambda: unet_tiny(1, 1),
    "unet_tiny_tc_gemm": lambda: unet_tiny(0, 1),
    "unet_tiny_tc": lambda: unet_tiny(0, 0),
    "unet_tiny_tc_b2": lambda: unet_tiny(0, 0, B=2, Fr=2, H=16, W=32),
    "unet_tiny_pair": lambda: unet_tiny(2, 0),
    "unet_tiny_pair256": lambda: unet_tiny(3, 0, cfg_over=dict(block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4))),
    "unet_tiny_tc_fmha2": lambda: unet_tiny(0, 2),
    "svd_steps_simt": lambda: svd_steps(gemm_impl=1, attn_impl=1),
    "svd_steps_tc": lambda: svd_steps(),
    "svd_steps_tc_cfg": lambda: svd_steps(cfg_scale=3.0),
    "svd_steps_tc_graph": lambda: svd_steps(graph=True),
    # the up blocks' GroupNorm eps at its other candidate value (UNVERIFIED U1 in oracle/unet_torch.py)
    "unet_tiny_tc_up_eps_1e5": lambda: unet_tiny(0, 0, cfg_over=dict(norm_eps={"up": 1e-5})),
    "unet_tiny_tc_fmha4": lambda: unet_tiny(0, 4),
    # the same through the Python orchestration (ctypes launch per kernel); the default above is csrc/unet.cu
    "unet_tiny_tc_pyorch": lambda: unet_tiny(0, 0, orchestrator="python"),
    "unet_tiny_pair256_pyorch": lambda: unet_tiny(3, 0, orchestrator="python", cfg_over=dict(
        block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4))),
    "svd_steps_tc_cfg_pyorch": lambda: svd_steps(cfg_scale=3.0, orchestrator="python"),
    "svd_steps_tc_graph_pyorch": lambda: svd_steps(graph=True, orchestrator="python"),
}
# ragged shapes: frame counts and latent sizes whose token counts are no multiple of any tile (M tails at every level,
# 1 frame = the temporal attention over a single key, 64 x 72 -> S = 4608 runs the ping-pong FMHA (impl 7) inside the UNet)
UNET_CHECKS["unet_tiny_ragged_f5_24x40"] = lambda: unet_tiny(0, None, Fr=5, H=24, W=40)
UNET_CHECKS["unet_tiny_ragged_f1_8x24"] = lambda: unet_tiny(0, None, Fr=1, H=8, W=24)
UNET_CHECKS["unet_tiny_ragged_b2_f7_40x8"] = lambda: unet_tiny(0, None, B=2, Fr=7, H=40, W=8)
UNET_CHECKS["unet_tiny_auto_impl_64x72"] = lambda: unet_tiny(3, None, Fr=2, H=64, W=72, cfg_over=dict(
    block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4)))
UNET_CHECKS["unet_tiny_tc_fmha7"] = lambda: unet_tiny(0, 7)
UNET_CHECKS["svd_steps_ragged_cfg"] = lambda: svd_steps(cfg_scale=2.5, Fr=5, H=24, W=40)
UNET_CHECKS["clip_vision_tiny"] = lambda: clip_vision()
UNET_CHECKS["clip_vision_hd80"] = lambda: clip_vision(B=1, hidden_size=320, num_attention_heads=4, intermediate_size=1280,
                                                      num_hidden_layers=3, image_size=224, projection_dim=128)
UNET_CHECKS["clip_vision_hd80_gemm_attention"] = lambda: clip_vision(B=1, hidden_size=320, num_attention_heads=4, intermediate_size=1280,
                                                                     num_hidden_layers=2, image_size=224, projection_dim=128, attn="gemm")
UNET_CHECKS["clip_vision_tiny_graph"] = lambda: clip_vision(use_graph=True)
UNET_CHECKS["vae_decode_tiny"] = lambda: vae_decode()
UNET_CHECKS["vae_decode_tiny_b2"] = lambda: vae_decode(B=2, Fr=2, h=16, w=8)
UNET_CHECKS["vae_decode_3level"] = lambda: vae_decode(B=1, Fr=2, h=8, w=16, block_out_channels=(64, 128, 256), layers_per_block=2)
UNET_CHECKS["vae_encode_tiny"] = lambda: vae_encode()
UNET_CHECKS["vae_encode_3level"] = lambda: vae_encode(N=1, H=128, W=128, block_out_channels=(64, 128, 256))
# the transformer feed-forwards through the fused kernel (svdpp_ff_geglu_f16; "ff_fused", off by default)
UNET_CHECKS["unet_tiny_tc_ff_fused"] = _tuned(UNET_CHECKS["unet_tiny_tc"], ff_fused=1)
UNET_CHECKS["unet_tiny_ragged_ff_fused"] = _tuned(UNET_CHECKS["unet_tiny_ragged_b2_f7_40x8"], ff_fused=1)
UNET_CHECKS["svd_steps_tc_cfg_ff_fused"] = _tuned(UNET_CHECKS["svd_steps_tc_cfg"], ff_fused=1)
UNET_CHECKS["unet_tiny_tc_copyout"] = _tuned(UNET_CHECKS["unet_tiny_tc"], tma_store=0)
UNET_CHECKS["unet_tiny_tc_zigzag"] = _tuned(UNET_CHECKS["unet_tiny_tc"], zigzag=1)      # producer -> consumer direction flips
UNET_CHECKS["unet_tiny_pair256_zigzag"] = _tuned(UNET_CHECKS["unet_tiny_pair256"], zigzag=1)
UNET_CHECKS["svd_steps_tc_cfg_zigzag"] = _tuned(UNET_CHECKS["svd_steps_tc_cfg"], zigzag=1)
UNET_CHECKS["unet_tiny_pair256_pdl"] = _tuned(UNET_CHECKS["unet_tiny_pair256"], pdl=1)
UNET_CHECKS["unet_tiny_pair256_r1ldg"] = _tuned(UNET_CHECKS["unet_tiny_pair256"], tma_r1=0)
UNET_CHECKS["unet_tiny_tc_dma"] = _tuned(UNET_CHECKS["unet_tiny_tc"], epi_dma=2)
UNET_CHECKS["unet_tiny_pair256_dma"] = _tuned(UNET_CHECKS["unet_tiny_pair256"], epi_dma=2)
UNET_CHECKS["unet_tiny_pair256_nodma"] = _tuned(UNET_CHECKS["unet_tiny_pair256"], epi_dma=0)
UNET_CHECKS["svd_steps_tc_graph_pdl"] = _tuned(UNET_CHECKS["svd_steps_tc_graph"], pdl=1)
