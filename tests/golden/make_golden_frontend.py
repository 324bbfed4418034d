"""Golden fixtures for the image -> video front / back end, generated FROM THE REFERENCE SCRIPT ITSELF.

Run in the build container (where /root/reference exists):   python tests/golden/make_golden_frontend.py
Nothing here is needed at test time; tests/test_frontend_host.py only reads the files this script wrote.

What is pinned and how (reference ``scripts/generate_video_demo.py``, imported unmodified from /root/reference):
  generate_cli.json   ``parse_args`` defaults (:33-59); SHA-256 of ``load_and_preprocess_image`` (:71-89) outputs for
                      three seeded images (wider, taller and exactly the target size)
  frontend.npz        ``encode_image`` (:92-152) and ``decode_latents`` (:154-195) driven on CPU / fp32 with a seeded
                      miniature CLIPVisionModelWithProjection (the real transformers class) and a miniature
                      AutoencoderKLTemporalDecoder (oracle/vae_torch.py): pins the conventions of
                      ``video-diffusion-pipeline-parallel_b200/frontend.py`` - pixel-space noise augmentation, ``mode()``, no
                      scaling factor on the image latents, ``1 / scaling_factor`` and per-chunk ``num_frames`` on decode
"""
import hashlib
import importlib.util
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_SCRIPT = "/root/reference/scripts/generate_video_demo.py"

TINY_CLIP = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=2, image_size=28,
                 patch_size=14, projection_dim=32, hidden_act="gelu")
TINY_VAE = dict(block_out_channels=(32, 32), layers_per_block=1)
CROP_CASES = [(300, 200, 64, 96), (100, 400, 64, 96), (96, 64, 64, 96)]      # (src_w, src_h, target_h, target_w)


def seeded_image(w: int, h: int, seed: int):
    from PIL import Image
    rng = np.random.default_rng(seed)
    return Image.fromarray(rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8), "RGB")


def state_checksum(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def build_modules():
    """The seeded miniature modules both sides of the test use (fp32, CPU)."""
    from transformers import CLIPImageProcessor, CLIPVisionConfig, CLIPVisionModelWithProjection
    sys.path.insert(0, ROOT)
    from oracle.vae_torch import AutoencoderKLTemporalDecoder
    torch.manual_seed(11)
    clip = CLIPVisionModelWithProjection(CLIPVisionConfig(**TINY_CLIP)).eval()
    torch.manual_seed(12)
    vae = AutoencoderKLTemporalDecoder(**TINY_VAE, force_upcast=False).eval()
    fx = CLIPImageProcessor(size={"shortest_edge": TINY_CLIP["image_size"]},
                            crop_size={"height": TINY_CLIP["image_size"], "width": TINY_CLIP["image_size"]})
    return clip, vae, fx


def main() -> None:
    spec = importlib.util.spec_from_file_location("ref_generate_video_demo", REF_SCRIPT)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)

    # ---- command line + centre crop
    argv = sys.argv
    sys.argv = ["generate_video_demo.py", "--input-image", "x.png"]
    try:
        defaults = vars(ref.parse_args())
    finally:
        sys.argv = argv
    crops = []
    with tempfile.TemporaryDirectory() as tmp:
        for i, (w, h, th, tw) in enumerate(CROP_CASES):
            path = os.path.join(tmp, f"img{i}.png")
            seeded_image(w, h, 100 + i).save(path)
            out = ref.load_and_preprocess_image(path, th, tw)
            crops.append(dict(src=[w, h], target=[th, tw], seed=100 + i, size=list(out.size),
                              sha256=hashlib.sha256(np.asarray(out).tobytes()).hexdigest()))
    with open(os.path.join(HERE, "generate_cli.json"), "w") as fh:
        json.dump(dict(defaults=defaults, crops=crops), fh, indent=1, sort_keys=True)

    # ---- encode_image / decode_latents
    clip, vae, fx = build_modules()
    dev, dt = torch.device("cpu"), torch.float32
    image = seeded_image(96, 64, 7)
    out = {"clip_checksum": state_checksum(clip.state_dict()), "vae_checksum": state_checksum(vae.state_dict())}
    arrays = {}
    for tag, strength in (("aug", 0.02), ("noaug", 0.0)):
        torch.manual_seed(7)
        emb, lat = ref.encode_image(image, clip, fx, vae, dev, dt, num_frames=3, noise_aug_strength=strength)
        arrays[f"emb_{tag}"], arrays[f"lat_{tag}"] = emb.numpy(), lat.numpy()
    torch.manual_seed(8)
    z = torch.randn(1, 4, 5, 8, 12)
    arrays["z"] = z.numpy()
    for chunk in (2, 14):
        arrays[f"frames_chunk{chunk}"] = ref.decode_latents(z, vae, 5, decode_chunk_size=chunk).numpy()
    np.savez_compressed(os.path.join(HERE, "frontend.npz"), **arrays)
    with open(os.path.join(HERE, "frontend.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print({k: v.shape for k, v in arrays.items()}, out)


if __name__ == "__main__":
    main()
