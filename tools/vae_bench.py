"""Front / back end at full size (SVD: 576x1024 frames, latent 72x128): NativeVAE decode + encode and NativeCLIPVision
(ViT-H/14) on the native kernels, against the torch restatement / the transformers class on torch's library kernels
(the reference decodes in fp32: force_upcast).  Random-init weights of the real architectures.
   python tools/vae_bench.py [--frames 25] [--chunk 14] [--no-library] [--clip-only]   ->  gpurun_out/vae_bench.json"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vdpp_b200  # noqa: E402,F401
from oracle.vae_torch import AutoencoderKLTemporalDecoder  # noqa: E402
from vdpp_b200 import native  # noqa: E402
from vdpp_b200.frontend import decode_latents  # noqa: E402
from vdpp_b200.models.native_clip import NativeCLIPVision  # noqa: E402
from vdpp_b200.models.native_vae import NativeVAE  # noqa: E402


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), out


def diff(a, b):
    a, b = a.float(), b.float()
    return dict(max_abs=(a - b).abs().max().item(), ref_absmax=b.abs().max().item(),
                cos=torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0).item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=25)
    ap.add_argument("--chunk", type=int, default=14)
    ap.add_argument("--no-library", action="store_true")
    ap.add_argument("--clip-only", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    res = {"frames": a.frames, "chunk": a.chunk}
    if not a.clip_only:
        vae_part(a, dev, res)
    clip_part(dev, res)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "vae_bench.json"), "w"), indent=1)
    print(json.dumps(res))


def vae_part(a, dev, res):
    torch.manual_seed(0)
    lib = AutoencoderKLTemporalDecoder().to(dev).half().eval()
    vae = NativeVAE(lib.state_dict(), device=dev)
    res["vae_weight_MB"] = vae.weight_bytes() / 1e6
    lat = torch.randn(1, 4, a.frames, 72, 128, device=dev).half()
    l0 = native.LAUNCHES
    ms, frames = timed(lambda: decode_latents(lat, vae, a.frames, a.chunk))
    res["decode_native_ms"] = ms
    res["decode_launches"] = (native.LAUNCHES - l0) // 4
    res["decode_finite"] = bool(torch.isfinite(frames).all())
    res["frames_shape"] = list(frames.shape)
    img = torch.rand(1, 3, 576, 1024, device=dev).half() * 2 - 1
    ms, z = timed(lambda: vae.encode(img).latent_dist.mode())
    res["encode_native_ms"] = ms
    res["encode_finite"] = bool(torch.isfinite(z).all())
    if not a.no_library:
        # the reference's path: fp32 VAE (force_upcast), torch library kernels, same chunking
        lib32 = lib.float()
        t0 = time.time()
        ms_lib, frames_lib = timed(lambda: decode_latents(lat.float(), lib32, a.frames, a.chunk), n=1)
        res["decode_library_fp32_ms"] = ms_lib
        res["decode_native_vs_library_fp32"] = diff(frames, frames_lib)
        ms_lib, z_lib = timed(lambda: lib32.encode(img.float()).latent_dist.mode(), n=1)
        res["encode_library_fp32_ms"] = ms_lib
        res["encode_native_vs_library_fp32"] = diff(z, z_lib)
        res["library_wall_s"] = time.time() - t0
        del lib32, frames_lib
    del lib, vae, frames
    torch.cuda.empty_cache()


def clip_part(dev, res):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    cfg = dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=16, image_size=224,
               patch_size=14, projection_dim=1024, hidden_act="gelu")
    torch.manual_seed(1)
    clip_lib = CLIPVisionModelWithProjection(CLIPVisionConfig(**cfg)).to(dev).half().eval()
    clip = NativeCLIPVision(clip_lib.state_dict(), config=cfg, device=dev)
    px = torch.randn(1, 3, 224, 224, device=dev).half()
    l0 = native.LAUNCHES
    ms, emb = timed(lambda: clip(px).image_embeds)
    res["clip_native_ms"] = ms
    res["clip_launches"] = (native.LAUNCHES - l0) // 4
    clip_g = NativeCLIPVision(clip_lib.state_dict(), config=cfg, device=dev, use_graph=True)
    ms_g, emb_g = timed(lambda: clip_g(px).image_embeds)
    res["clip_native_graph_ms"] = ms_g
    res["clip_graph_equals_eager"] = bool(torch.equal(emb_g, emb))
    del clip_g
    clip_old = NativeCLIPVision(clip_lib.state_dict(), config=cfg, device=dev, attn="gemm")
    ms_old, emb_old = timed(lambda: clip_old(px).image_embeds)
    res["clip_native_gemm_attention_ms"] = ms_old
    res["clip_fused_vs_gemm_attention"] = diff(emb, emb_old)
    del clip_old
    with torch.no_grad():
        ms_lib, emb_lib = timed(lambda: clip_lib(pixel_values=px).image_embeds)
        emb32 = clip_lib.float()(pixel_values=px.float()).image_embeds
    res["clip_library_fp16_ms"] = ms_lib
    res["clip_native_vs_library_fp32"] = diff(emb, emb32)
    res["clip_library_fp16_vs_fp32"] = diff(emb_lib, emb32)


if __name__ == "__main__":
    main()
