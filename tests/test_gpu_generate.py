"""GPU tests of the image -> video run (``modes/generate_video.py``, the reference's ``scripts/generate_video_demo.py``):
miniature CLIP / VAE / UNet of the real architectures built by the product's own loaders (``from_pretrained`` with
``random-init`` and with a local snapshot directory), the whole run against its hand-made composition, files written."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

TINY_UNET = dict(block_out_channels=(64, 128, 128, 128), num_attention_heads=(1, 2, 2, 2), num_frames=3, cross_attention_dim=1024)
TINY_VAE = dict(block_out_channels=(64, 64, 128, 128), layers_per_block=1)
TINY_CLIP = dict(hidden_size=192, intermediate_size=384, num_hidden_layers=2, num_attention_heads=4, image_size=56,
                 patch_size=14, projection_dim=1024, hidden_act="gelu")


def _extractor():
    from transformers import CLIPImageProcessor
    return CLIPImageProcessor(size={"shortest_edge": 56}, crop_size={"height": 56, "width": 56})


def _networks(seed=1, steps=4):
    from vdpp_b200.models import StableVideoUNet
    from vdpp_b200.models.native_clip import NativeCLIPVision
    from vdpp_b200.models.native_vae import NativeVAE
    mid = f"random-init:{seed}"
    clip = NativeCLIPVision.from_pretrained(mid, config=TINY_CLIP)
    vae = NativeVAE.from_pretrained(mid, config=TINY_VAE)
    model = StableVideoUNet.from_pretrained(mid, timesteps=StableVideoUNet._default_timestep_schedule(steps), config=TINY_UNET)
    return clip, vae, model


def _args(tmp_path, *extra):
    from vdpp_b200.modes import generate_video as gv
    return gv.build_parser().parse_args(["--input-image", "synthetic:1", "--height", "128", "--width", "128", "--num-frames", "3",
                                         "--total-steps", "4", "--num-samples", "2", "--seed", "9", "--output-dir", str(tmp_path),
                                         *extra])


@pytest.mark.parametrize("guidance", ["3.0", "1.0"])
def test_generate_video_equals_its_composition(tmp_path, capsys, guidance):
    from PIL import Image
    from vdpp_b200 import frontend, native
    from vdpp_b200.modes import generate_video as gv
    dev = torch.device("cuda", 0)
    clip, vae, model = _networks()
    args = _args(tmp_path, "--guidance-scale", guidance, "--decode-chunk-size", "2", "--save-frames")
    l0 = native.LAUNCHES
    rec = gv.generate(args, image_encoder=clip, vae=vae, model=model, feature_extractor=_extractor())
    assert native.LAUNCHES > l0
    frames = rec["frames"]
    assert len(frames) == 2 and all(tuple(f.shape) == (1, 3, 3, 128, 128) and f.dtype == torch.float32 for f in frames)
    assert rec["frames_finite"] and not torch.equal(frames[0], frames[1])
    line = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("GENERATE_JSON=")]
    assert len(line) == 1
    printed = json.loads(line[0][len("GENERATE_JSON="):])
    assert printed["diffusion_s"] > 0 and printed["decode_s"] > 0 and len(printed["files"]) == 2
    import numpy as np
    for k, path in enumerate(printed["files"]):
        with Image.open(path) as g:
            assert g.n_frames == 3 and g.size == (128, 128)
            # quantised on the GPU into the fixed colour cube: every pixel within one cube step of the RGB bytes
            g.seek(1)
            back = np.asarray(g.convert("RGB"), dtype=np.int32)
        rgb = gv.frames_to_uint8(frames[k])
        assert rgb.shape == (3, 128, 128, 3) and rgb.dtype == np.uint8
        assert np.abs(back - rgb[1].astype(np.int32)).max() <= 52
        want = ((frames[k][0].permute(1, 2, 3, 0) + 1) / 2 * 255).clamp(0, 255).to(torch.uint8).cpu().numpy()
        assert np.array_equal(rgb, want)              # the reference's expression, bit for bit
    assert len([f for f in os.listdir(tmp_path) if f.endswith(".png")]) == 6

    # the same run put together by hand from the public pieces
    image = gv.load_and_preprocess_image("synthetic:1", 128, 128)
    gen = torch.Generator(device=dev).manual_seed(9)
    emb, lat = frontend.encode_image(image, clip, _extractor(), vae, dev, torch.float16, 3, 0.02, generator=gen)
    assert tuple(emb.shape) == (1, 1, 1024) and tuple(lat.shape) == (1, 4, 3, 16, 16)
    gs = float(guidance) if float(guidance) > 1.0 else None
    model.set_conditioning(emb, lat, fps=7, motion_bucket_id=127, noise_aug_strength=0.02, guidance_scale=gs, num_frames=3)
    for idx in range(2):
        torch.manual_seed(9 + idx)
        x = torch.randn(1, 4, 3, 16, 16, device=dev, dtype=torch.float16) * model.init_noise_sigma
        for s in range(4):
            x = model(x, s)
        want = frontend.decode_latents(x, vae, 3, decode_chunk_size=2)
        assert torch.equal(frames[idx], want)


def test_loaders_read_a_local_snapshot_like_the_hub_layout(tmp_path):
    """``from_pretrained(<dir>)`` of all three networks on a snapshot written in the hub layout gives the same run as
    the state dicts it was written from."""
    from safetensors.torch import save_file
    from vdpp_b200.models import StableVideoUNet, frontend_weights as fw
    from vdpp_b200.models.native_clip import NativeCLIPVision
    from vdpp_b200.models.native_vae import NativeVAE
    from vdpp_b200.models.svd_weights import random_state_dict
    snap = tmp_path / "snap"
    sds = dict(vae=fw.random_state_dict(fw.vae_param_shapes(TINY_VAE), seed=5, device="cuda"),
               image_encoder=fw.random_state_dict(fw.clip_param_shapes(TINY_CLIP), seed=6, device="cuda"),
               unet=random_state_dict(TINY_UNET, seed=7, device="cuda"))
    names = dict(vae="diffusion_pytorch_model.fp16.safetensors", image_encoder="model.fp16.safetensors",
                 unet="diffusion_pytorch_model.fp16.safetensors")
    for sub, sd in sds.items():
        (snap / sub).mkdir(parents=True)
        save_file({k: v.cpu().contiguous() for k, v in sd.items()}, str(snap / sub / names[sub]))
    json.dump(dict(TINY_VAE, force_upcast=True), open(snap / "vae" / "config.json", "w"))
    json.dump(TINY_CLIP, open(snap / "image_encoder" / "config.json", "w"))
    vae_a, vae_b = NativeVAE.from_pretrained(str(snap)), NativeVAE(sds["vae"], config=TINY_VAE)
    z = torch.randn(3, 4, 8, 8, device="cuda").half()
    assert torch.equal(vae_a.decode(z, num_frames=3).sample, vae_b.decode(z, num_frames=3).sample)
    assert vae_a.config.force_upcast is False            # fp32 accumulation inside the kernels: nothing to up-cast
    clip_a, clip_b = NativeCLIPVision.from_pretrained(str(snap)), NativeCLIPVision(sds["image_encoder"], config=TINY_CLIP)
    px = torch.randn(1, 3, 56, 56, device="cuda").half()
    assert torch.equal(clip_a(px).image_embeds, clip_b(px).image_embeds)
    ts = StableVideoUNet._default_timestep_schedule(4)
    m = StableVideoUNet.from_pretrained(str(snap), timesteps=ts, config=TINY_UNET)
    m.set_dummy_conditioning(1, 3, 16, 16, torch.device("cuda"))
    assert torch.isfinite(m(torch.randn(1, 4, 3, 16, 16, device="cuda").half() * m.init_noise_sigma, 0)).all()
    with pytest.raises(FileNotFoundError):
        NativeVAE.from_pretrained("stabilityai/stable-video-diffusion-img2vid-xt")


def test_random_init_vae_and_clip_agree_with_the_library_modules():
    """The generated state dicts drive the native modules to the same function as the torch restatement (VAE) and the
    real transformers class (CLIP) loaded with the same tensors - every key of the inventory is consumed correctly."""
    from oracle.vae_torch import AutoencoderKLTemporalDecoder
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    from vdpp_b200.models import frontend_weights as fw
    from vdpp_b200.models.native_clip import NativeCLIPVision
    from vdpp_b200.models.native_vae import NativeVAE
    sd = fw.random_state_dict(fw.vae_param_shapes(TINY_VAE), seed=2, device="cuda")
    lib = AutoencoderKLTemporalDecoder(**TINY_VAE).cuda().float().eval()
    lib.load_state_dict({k: v.float() for k, v in sd.items()}, strict=True)
    vae = NativeVAE(sd, config=TINY_VAE)
    z = torch.randn(3, 4, 8, 8, device="cuda").half()
    with torch.no_grad():
        want = lib.decode(z.float(), num_frames=3).sample
    got = vae.decode(z, num_frames=3).sample.float()
    assert (got - want).abs().max().item() <= 2e-2 * want.abs().max().item() + 2e-3
    csd = fw.random_state_dict(fw.clip_param_shapes(TINY_CLIP), seed=3, device="cuda")
    clib = CLIPVisionModelWithProjection(CLIPVisionConfig(**TINY_CLIP)).cuda().float().eval()
    clib.load_state_dict({k: v.float() for k, v in csd.items()}, strict=True)
    px = torch.randn(2, 3, 56, 56, device="cuda").half()
    with torch.no_grad():
        cwant = clib(pixel_values=px.float()).image_embeds
    cgot = NativeCLIPVision(csd, config=TINY_CLIP)(px).image_embeds.float()
    assert (cgot - cwant).abs().max().item() <= 1e-2 * cwant.abs().max().item() + 3e-3
