"""Front / back end of the image -> video run, host side (CPU): the conventions of ``frontend.py`` against outputs recorded
from the reference script's own functions (tests/golden/make_golden_frontend.py), the generation CLI against the
reference's flags, the VAE / CLIP parameter inventories against the torch restatement / the real transformers class,
and the local-checkpoint reader."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "golden"))

import vdpp_b200  # noqa: E402,F401
from vdpp_b200 import frontend  # noqa: E402
from vdpp_b200.models import frontend_weights as fw  # noqa: E402
from vdpp_b200.modes import generate_video as gv  # noqa: E402

import make_golden_frontend as mg  # noqa: E402  (builders of the seeded miniature modules; reads nothing from /root/reference)

GOLD = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def modules():
    clip, vae, fx = mg.build_modules()
    want = json.load(open(os.path.join(GOLD, "frontend.json")))
    assert mg.state_checksum(clip.state_dict()) == want["clip_checksum"], "seeded CLIP init changed: regenerate the golden"
    assert mg.state_checksum(vae.state_dict()) == want["vae_checksum"], "seeded VAE init changed: regenerate the golden"
    return clip, vae, fx


@pytest.mark.parametrize("tag,strength", [("aug", 0.02), ("noaug", 0.0)])
def test_encode_image_matches_reference_function(modules, tag, strength):
    clip, vae, fx = modules
    gold = np.load(os.path.join(GOLD, "frontend.npz"))
    torch.manual_seed(7)       # generator=None: the global stream, as the reference's randn_like
    emb, lat = frontend.encode_image(mg.seeded_image(96, 64, 7), clip, fx, vae, torch.device("cpu"), torch.float32,
                                     num_frames=3, noise_aug_strength=strength)
    assert tuple(emb.shape) == gold[f"emb_{tag}"].shape and tuple(lat.shape) == gold[f"lat_{tag}"].shape
    np.testing.assert_allclose(emb.numpy(), gold[f"emb_{tag}"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(lat.numpy(), gold[f"lat_{tag}"], rtol=0, atol=1e-6)
    # the frames are copies of one image latent
    assert torch.equal(lat[:, :, 0], lat[:, :, 2])


def test_encode_image_accepts_a_tensor_and_a_generator(modules):
    clip, vae, fx = modules
    img = mg.seeded_image(96, 64, 7)
    img01 = torch.from_numpy(np.asarray(img, dtype=np.float32) / 255.0).permute(2, 0, 1)[None]
    px = fx(images=img, return_tensors="pt").pixel_values
    emb_t, lat_t = frontend.encode_image(img01, clip, lambda x: px, vae, torch.device("cpu"), torch.float32, 3, 0.0)
    emb_p, lat_p = frontend.encode_image(img, clip, fx, vae, torch.device("cpu"), torch.float32, 3, 0.0)
    assert torch.allclose(emb_t, emb_p, atol=1e-6) and torch.allclose(lat_t, lat_p, atol=1e-6)
    g1, g2 = torch.Generator().manual_seed(5), torch.Generator().manual_seed(5)
    a = frontend.encode_image(img, clip, fx, vae, torch.device("cpu"), torch.float32, 3, 0.02, generator=g1)[1]
    b = frontend.encode_image(img, clip, fx, vae, torch.device("cpu"), torch.float32, 3, 0.02, generator=g2)[1]
    assert torch.equal(a, b) and not torch.equal(a, lat_p)


@pytest.mark.parametrize("chunk", [2, 14])
def test_decode_latents_matches_reference_function(modules, chunk):
    _, vae, _ = modules
    gold = np.load(os.path.join(GOLD, "frontend.npz"))
    frames = frontend.decode_latents(torch.from_numpy(gold["z"]), vae, 5, decode_chunk_size=chunk)
    assert frames.dtype == torch.float32 and tuple(frames.shape) == gold[f"frames_chunk{chunk}"].shape
    np.testing.assert_allclose(frames.numpy(), gold[f"frames_chunk{chunk}"], rtol=0, atol=1e-6)


def test_chunking_changes_the_temporal_mix(modules):
    """The temporal decoder mixes the frames of a chunk: the two chunkings are different functions, so the default of the
    generation CLI must stay the reference's (4)."""
    gold = np.load(os.path.join(GOLD, "frontend.npz"))
    assert np.abs(gold["frames_chunk2"] - gold["frames_chunk14"]).max() > 1e-4
    assert gv.build_parser().parse_args(["--input-image", "x"]).decode_chunk_size == 4


# ---------------------------------------------------------------------------------------------- CLI
def test_cli_flags_and_defaults_are_the_references():
    want = json.load(open(os.path.join(GOLD, "generate_cli.json")))["defaults"]
    got = vars(gv.build_parser().parse_args(["--input-image", "x.png"]))
    for key, value in want.items():
        assert key in got, key
        assert got[key] == value, (key, got[key], value)
    with pytest.raises(SystemExit):
        gv.build_parser().parse_args([])            # --input-image is required, as in the reference


def test_center_crop_matches_reference_function(tmp_path):
    for case in json.load(open(os.path.join(GOLD, "generate_cli.json")))["crops"]:
        (w, h), (th, tw) = case["src"], case["target"]
        path = str(tmp_path / "img.png")
        mg.seeded_image(w, h, case["seed"]).save(path)
        out = gv.load_and_preprocess_image(path, th, tw)
        assert list(out.size) == case["size"] == [tw, th]
        assert hashlib.sha256(np.asarray(out).tobytes()).hexdigest() == case["sha256"], case


def test_synthetic_image_and_gif_round_trip(tmp_path):
    from PIL import Image
    img = gv.load_and_preprocess_image("synthetic:3", 64, 96)
    assert img.size == (96, 64) and img.mode == "RGB"
    assert np.asarray(img).std() > 10
    assert np.array_equal(np.asarray(img), np.asarray(gv.synthetic_image(64, 96, 3)))
    frames = torch.rand(1, 3, 5, 16, 24) * 2 - 1
    u8 = gv.frames_to_uint8(frames)
    assert u8.shape == (5, 16, 24, 3) and u8.dtype == np.uint8
    want0 = ((frames[0, :, 0].permute(1, 2, 0) + 1) / 2 * 255).clamp(0, 255).to(torch.uint8).numpy()
    assert np.array_equal(u8[0], want0)
    path = str(tmp_path / "v.gif")
    gv.save_gif(frames, path, fps=7)
    with Image.open(path) as g:
        assert g.n_frames == 5 and g.size == (24, 16)


def test_generate_needs_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no GPU"):
        gv.main(["--input-image", "synthetic", "--no-files"])


# ---------------------------------------------------------------------------------------------- weights
def test_vae_inventory_matches_the_restatement():
    from oracle.vae_torch import AutoencoderKLTemporalDecoder, tiny_vae_config
    for cfg in (None, tiny_vae_config(), dict(block_out_channels=(64, 128, 128), layers_per_block=2)):
        ref = {k: tuple(v.shape) for k, v in AutoencoderKLTemporalDecoder(**(cfg or {})).state_dict().items()}
        assert dict(fw.vae_param_shapes(cfg)) == ref
    assert fw.param_count(fw.vae_param_shapes()) == 97_742_847          # the published SVD VAE size


def test_clip_inventory_matches_transformers():
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    ref = {k: tuple(v.shape) for k, v in CLIPVisionModelWithProjection(CLIPVisionConfig(**mg.TINY_CLIP)).state_dict().items()}
    assert dict(fw.clip_param_shapes(mg.TINY_CLIP)) == ref
    assert fw.param_count(fw.clip_param_shapes()) == 632_076_800         # CLIP ViT-H/14 vision tower + projection


def test_random_init_is_seeded_and_loads_into_the_library_modules():
    from oracle.vae_torch import AutoencoderKLTemporalDecoder
    shapes = fw.vae_param_shapes(mg.TINY_VAE)
    a = fw.random_state_dict(shapes, seed=3, device="cpu", dtype=torch.float32)
    b = fw.random_state_dict(shapes, seed=3, device="cpu", dtype=torch.float32)
    c = fw.random_state_dict(shapes, seed=4, device="cpu", dtype=torch.float32)
    assert all(torch.equal(a[k], b[k]) for k in a) and any(not torch.equal(a[k], c[k]) for k in a)
    AutoencoderKLTemporalDecoder(**mg.TINY_VAE).load_state_dict(a, strict=True)
    assert float(a["decoder.mid_block.resnets.0.time_mixer.mix_factor"]) == 0.5
    assert torch.equal(a["decoder.conv_norm_out.weight"], torch.ones(32))


def test_local_snapshot_reader(tmp_path):
    from safetensors.torch import save_file
    sd = fw.random_state_dict(fw.vae_param_shapes(mg.TINY_VAE), seed=1, device="cpu", dtype=torch.float16)
    vdir = tmp_path / "snap" / "vae"
    vdir.mkdir(parents=True)
    save_file({k: v.float() for k, v in sd.items()}, str(vdir / "diffusion_pytorch_model.safetensors"))
    save_file(sd, str(vdir / "diffusion_pytorch_model.fp16.safetensors"))
    json.dump(dict(mg.TINY_VAE, force_upcast=True, _class_name="AutoencoderKLTemporalDecoder"), open(vdir / "config.json", "w"))
    got, cfg = fw.load_component(str(tmp_path / "snap"), "vae", device="cpu")
    assert set(got) == set(sd) and all(got[k].dtype == torch.float16 and torch.equal(got[k], sd[k]) for k in sd)
    assert cfg["layers_per_block"] == 1 and list(cfg["block_out_channels"]) == [32, 32]
    got2, _ = fw.load_component(str(vdir), "vae", device="cpu")         # the component folder itself
    assert set(got2) == set(sd)
    with pytest.raises(FileNotFoundError, match="no network"):
        fw.load_component("stabilityai/stable-video-diffusion-img2vid-xt", "vae", device="cpu")
    with pytest.raises(FileNotFoundError, match="no .safetensors"):
        (tmp_path / "empty" / "vae").mkdir(parents=True)
        fw.load_component(str(tmp_path / "empty"), "vae", device="cpu")
    assert fw.is_random_init("random-init:7") and fw.random_init_seed("random-init:7") == 7 and fw.random_init_seed("random-init") == 0


def test_colour_cube_palette_and_index_restatement():
    """The fixed palette of svdpp_frames_to_bytes and the numpy restatement of its quantiser (the oracle of the
    frames_to_bytes_* GPU checks): 252 distinct colours, rounding to the nearest level without dither (error at most half a
    cube step), within one step with the ordered dither, and a dither that is unbiased over a 4 x 4 cell."""
    import kernel_checks as kh          # importable without a GPU: only its check functions touch the device
    from vdpp_b200 import native
    pal = np.array(native.cube_palette(), dtype=np.int32).reshape(256, 3)
    assert len({tuple(c) for c in pal[:252]}) == 252 and (pal[252:] == 0).all()
    assert tuple(pal[0]) == (0, 0, 0) and tuple(pal[251]) == (255, 255, 255) and tuple(pal[1 * 42 + 2 * 6 + 3]) == (51, 85, 153)
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 256, size=(2, 8, 12, 3), dtype=np.uint8)
    rgb[0, 0, :4] = [[0, 0, 0], [255, 255, 255], [25, 21, 26], [26, 22, 25]]
    plain = kh.cube_index_reference(rgb, dither=False)
    err = np.abs(pal[plain] - rgb.astype(np.int32))
    assert err[..., 0].max() <= 26 and err[..., 1].max() <= 22 and err[..., 2].max() <= 26
    dith = kh.cube_index_reference(rgb, dither=True)
    errd = np.abs(pal[dith] - rgb.astype(np.int32))
    assert errd[..., 0].max() <= 51 and errd[..., 1].max() <= 43 and errd[..., 2].max() <= 51 and int(dith.max()) < 252
    flat = np.full((1, 4, 4, 3), 100, dtype=np.uint8)                # a flat grey: the 16 dither thresholds average it out
    mean = pal[kh.cube_index_reference(flat, dither=True)].reshape(-1, 3).mean(axis=0)
    assert np.abs(mean - 100).max() <= 4
