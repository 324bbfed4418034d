"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container (where /root/reference exists):   python tests/golden/make_golden.py
Nothing here is needed at test time; the tests only read the .json / .npz files this script wrote.

What is pinned and how:
  step_assignment.json   assign_steps(T, W, r) for a grid of (T, W), straight from
                         /root/reference/src/pipeline/step_assignment.py (incl. which inputs raise)
  dummy_unet.npz         reference DummyUNet (seeded weights, saved) forward outputs at three steps
  svd_step.npz           the UNMODIFIED reference StableVideoUNet wrapper (src/models/svd_unet.py) driven
                         on CPU/fp32 with a miniature UNet (oracle/unet_torch.py via oracle/diffusers_shim),
                         with and without classifier-free guidance: pins the wrapper arithmetic
                         (scale, cat, permute, CFG, v-prediction Euler) of oracle/svd_step.py
  scheduler.json         known answers for the Karras table recorded in SURVEY.md section 8c and the
                         reference's EXPERIMENT_RESULTS.md:242-243
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

GOLDEN_UNET_CFG = dict(block_out_channels=(32, 64), down_attn=(True, False), num_attention_heads=(1, 1),
                       num_frames=3, cross_attention_dim=1024)


def state_checksum(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def main() -> None:
    sys.path.insert(0, ROOT)                                   # `oracle` package
    sys.path.insert(0, os.path.join(ROOT, "oracle", "diffusers_shim"))
    # import the reference's `src` package (not ours): put it first and make sure ours is not cached
    for m in [m for m in sys.modules if m == "src" or m.startswith("src.")]:
        del sys.modules[m]
    sys.path.insert(0, REF)
    from src.models.dummy_unet import DummyUNet as RefDummy
    from src.models.svd_unet import StableVideoUNet as RefWrapper
    from src.pipeline.step_assignment import assign_steps as ref_assign
    import src
    assert os.path.abspath(src.__file__).startswith(REF), src.__file__

    # ---- step assignment
    table = {}
    for T in (1, 7, 24, 25, 28, 29, 35, 105):
        for W in (1, 2, 3, 4, 5, 7, 8):
            rows = []
            for r in range(W):
                try:
                    sr = ref_assign(T, W, r)
                    rows.append([sr.start, sr.end])
                except ValueError:
                    rows.append("ValueError")
            table[f"{T}x{W}"] = rows
    errors = {}
    for name, args in dict(zero_steps=(0, 1, 0), neg_steps=(-1, 1, 0), zero_world=(28, 0, 0), rank_hi=(28, 4, 4),
                           rank_neg=(28, 4, -1)).items():
        try:
            ref_assign(*args)
            errors[name] = "ok"
        except ValueError:
            errors[name] = "ValueError"
    json.dump(dict(table=table, errors=errors), open(os.path.join(HERE, "step_assignment.json"), "w"), indent=1)

    # ---- DummyUNet
    torch.manual_seed(1234)
    m = RefDummy(channels=4, hidden_channels=16)
    torch.manual_seed(42)
    x = torch.randn(1, 4, 3, 8, 8)
    out = {}
    with torch.no_grad():
        for step in (0, 10, 27):
            out[f"out_step{step}"] = m(x, step).numpy()
    np.savez_compressed(os.path.join(HERE, "dummy_unet.npz"), x=x.numpy(),
                        **{"sd." + k: v.numpy() for k, v in m.state_dict().items()}, **out)

    # ---- reference wrapper around a miniature UNet (CPU, fp32)
    from oracle.unet_torch import UNetSpatioTemporalConditionModel
    torch.manual_seed(7)
    unet = UNetSpatioTemporalConditionModel(**GOLDEN_UNET_CFG).eval()
    chk = state_checksum(unet.state_dict())
    total = 25
    res = dict(unet_checksum=np.frombuffer(chk.encode(), dtype=np.uint8))
    for tag, gscale in (("nocfg", None), ("cfg", 3.0)):
        w = RefWrapper(unet=unet, timesteps=RefWrapper._default_timestep_schedule(total), dtype=torch.float32)
        torch.manual_seed(11)
        w.set_dummy_conditioning(batch_size=1, num_frames=3, height=8, width=8, device=torch.device("cpu"),
                                 guidance_scale=gscale)
        torch.manual_seed(42)
        lat = torch.randn(1, 4, 3, 8, 8) * w.init_noise_sigma
        res[f"{tag}_latent0"] = lat.numpy()
        x_ = lat
        for step in range(3):
            x_ = w(x_, step)
            res[f"{tag}_latent{step + 1}"] = x_.numpy()
        res[f"{tag}_sigmas"] = w.sigmas.numpy()
        res[f"{tag}_timesteps"] = w.scheduler_timesteps.numpy()
        res[f"{tag}_init_noise_sigma"] = np.array(w.init_noise_sigma)
        res[f"{tag}_added_time_ids"] = w._added_time_ids.numpy()
    res["default_schedule_25"] = np.array(RefWrapper._default_timestep_schedule(25))
    np.savez_compressed(os.path.join(HERE, "svd_step.npz"), **res)

    # ---- scheduler known answers (SURVEY.md 8c probe; EXPERIMENT_RESULTS.md:242-243)
    json.dump({
        "25": {"sigmas_head": [700.0, 545.7292, 421.5691], "sigmas_tail": [0.024803, 0.007882, 0.002, 0.0],
               "timestep_first": 1.63777, "timestep_last": -1.55365, "init_noise_sigma": 700.000732},
        "28": {"sigma_1": 561.2835}, "35": {"sigma_1": 587.731},
        "source": "SURVEY.md section 8c [probe]; reference EXPERIMENT_RESULTS.md:242-243 (sigmas[0]=700.0)",
    }, open(os.path.join(HERE, "scheduler.json"), "w"), indent=1)
    print("golden fixtures written; unet checksum", chk)


if __name__ == "__main__":
    main()
