"""Pin the oracle against the REAL diffusers package the moment it is importable.

``oracle/unet_torch.py`` and ``oracle/vae_torch.py`` restate ``UNetSpatioTemporalConditionModel`` and
``AutoencoderKLTemporalDecoder`` from memory (diffusers is neither installed nor vendored here), so every native-vs-
oracle parity test shares the oracle's guesses (the UNVERIFIED table in ``oracle/unet_torch.py``).  This script removes
the guesswork as soon as ``import diffusers`` works (a pip install on a networked box, or a wheel dropped into the image):

  1. builds the diffusers module with a small config and the default SVD config's block layout,
  2. checks that parameter names, shapes and the parameter count of the oracle equal the library's,
  3. copies the library's weights into the oracle and compares forward outputs in fp32 on CPU (tolerance 1e-5 relative),
  4. reads the GroupNorm eps of every block class out of the library module and compares it with ``NORM_EPS``,
  5. does the same for the scheduler tables (``EulerDiscreteScheduler`` with the reference's arguments,
     src/models/svd_unet.py:81-94) against ``oracle/scheduler.py``.

Exit code 0 = everything agrees; 1 = a difference (printed, with the name of the UNVERIFIED entry it settles);
2 = diffusers not importable (nothing could be verified - the state this repository was built in).
   python tools/verify_against_diffusers.py [--full]      (--full also builds the 1.5 B-parameter default config)
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    a = ap.parse_args()
    try:
        import diffusers  # noqa: F401
        from diffusers import AutoencoderKLTemporalDecoder as LibVAE
        from diffusers import EulerDiscreteScheduler
        from diffusers import UNetSpatioTemporalConditionModel as LibUNet
    except Exception as e:  # noqa: BLE001
        print(f"diffusers is not importable here ({type(e).__name__}: {e}); nothing verified")
        return 2
    import numpy as np
    import torch

    from oracle.scheduler import euler_karras_tables
    from oracle.unet_torch import NORM_EPS, UNetSpatioTemporalConditionModel
    from oracle.vae_torch import AutoencoderKLTemporalDecoder

    bad = []

    def check(cond, msg):
        print(("ok   " if cond else "DIFF ") + msg)
        if not cond:
            bad.append(msg)

    # ---- UNet: small config with every block type
    small = dict(block_out_channels=(64, 128, 128, 128), num_attention_heads=(1, 2, 2, 2), num_frames=3,
                 cross_attention_dim=1024)
    torch.manual_seed(0)
    lib = LibUNet(in_channels=8, out_channels=4, **small).eval()
    ora = UNetSpatioTemporalConditionModel(**small).eval()
    lsd, osd = lib.state_dict(), ora.state_dict()
    check(set(lsd) == set(osd), f"UNet state_dict keys ({len(lsd)} library / {len(osd)} oracle; "
                                f"only in library: {sorted(set(lsd) - set(osd))[:5]}, only in oracle: {sorted(set(osd) - set(lsd))[:5]})")
    check(all(tuple(lsd[k].shape) == tuple(osd[k].shape) for k in lsd if k in osd), "UNet parameter shapes")
    ora.load_state_dict({k: v for k, v in lsd.items() if k in osd}, strict=False)
    x = torch.randn(2, 3, 8, 16, 16)
    enc, ids = torch.randn(2, 1, 1024), torch.tensor([[5.0, 127.0, 0.02]] * 2)
    with torch.no_grad():
        yl = lib(x, torch.tensor(1.6377), enc, ids, return_dict=False)[0]
        yo = ora(x, torch.tensor(1.6377), enc, ids)[0]
    err = (yl - yo).abs().max().item() / max(yl.abs().max().item(), 1e-9)
    check(err < 1e-5, f"UNet forward, oracle vs diffusers, fp32: relative max error {err:.3e} "
                      "(settles U3 AlphaBlender, U4 GEGLU order, U5 sinusoid order, U6 first-frame context, U7 position add)")
    # eps per block class, read from the library module (settles U1 / U2)
    def eps_of(block):
        r = block.resnets[0]
        return r.spatial_res_block.norm1.eps, r.temporal_res_block.norm1.eps
    got = dict(down_attn=eps_of(lib.down_blocks[0])[0], down=eps_of(lib.down_blocks[-1])[0], mid=eps_of(lib.mid_block)[0],
               up=eps_of(lib.up_blocks[0])[0], up_attn=eps_of(lib.up_blocks[-1])[0],
               transformer=lib.down_blocks[0].attentions[0].norm.eps, out=lib.conv_norm_out.eps)
    print("library GroupNorm eps:", got)
    for k in ("down_attn", "down", "mid", "up", "transformer", "out"):
        check(abs(got[k] - NORM_EPS[k]) < 1e-12, f"NORM_EPS[{k!r}] = {NORM_EPS[k]} vs library {got[k]}  (U1 / U2)")
    check(abs(got["up_attn"] - NORM_EPS["up"]) < 1e-12, f"CrossAttnUpBlock eps {got['up_attn']} vs NORM_EPS['up'] {NORM_EPS['up']}")
    check(all(abs(a_ - b_) < 1e-12 for a_, b_ in [eps_of(lib.up_blocks[0])]), "temporal ResBlock eps equals the spatial one")
    if a.full:
        with torch.device("meta"):
            n_lib = sum(p.numel() for p in LibUNet().parameters())
            n_ora = sum(p.numel() for p in UNetSpatioTemporalConditionModel().parameters())
        check(n_lib == n_ora == 1_524_623_082, f"SVD UNet parameter count: library {n_lib}, oracle {n_ora}")

    # ---- VAE
    vsmall = dict(block_out_channels=(64, 128), layers_per_block=1)
    lv = LibVAE(**vsmall).eval()
    ov = AutoencoderKLTemporalDecoder(**vsmall).eval()
    lvs, ovs = lv.state_dict(), ov.state_dict()
    check(set(lvs) == set(ovs), f"VAE state_dict keys (only in library: {sorted(set(lvs) - set(ovs))[:5]}, "
                                f"only in oracle: {sorted(set(ovs) - set(lvs))[:5]})")
    ov.load_state_dict({k: v for k, v in lvs.items() if k in ovs}, strict=False)
    z, img = torch.randn(3, 4, 8, 16), torch.randn(1, 3, 32, 64)
    with torch.no_grad():
        dl = lv.decode(z, num_frames=3).sample
        do = ov.decode(z, num_frames=3).sample
        el = lv.encode(img).latent_dist.mode()
        eo = ov.encode(img).latent_dist.mode()
    check(((dl - do).abs().max() / dl.abs().max()).item() < 1e-5, "VAE decode, oracle vs diffusers, fp32 (V1, V2, V4, V6)")
    check(((el - eo).abs().max() / el.abs().max()).item() < 1e-5, "VAE encode, oracle vs diffusers, fp32 (V3, V5)")

    # ---- scheduler (reference svd_unet.py:81-94)
    for n in (25, 28, 35):
        s = EulerDiscreteScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                                   prediction_type="v_prediction", interpolation_type="linear", use_karras_sigmas=True,
                                   sigma_min=0.002, sigma_max=700.0, timestep_spacing="leading", timestep_type="continuous",
                                   steps_offset=1, rescale_betas_zero_snr=False)
        s.set_timesteps(n)
        sig, ts, ins = euler_karras_tables(n)
        check(np.allclose(s.sigmas.numpy(), sig.numpy(), rtol=1e-6, atol=1e-9), f"scheduler sigmas, n = {n}")
        check(np.allclose(s.timesteps.numpy(), ts.numpy(), rtol=1e-6, atol=1e-7), f"scheduler timesteps, n = {n}")
        check(abs(float(s.init_noise_sigma) - ins) < 1e-4, f"init_noise_sigma, n = {n}")
    print("ALL VERIFIED" if not bad else f"{len(bad)} difference(s): fix the oracle AND the native table, then re-run the parity tests")
    return 0 if not bad else 1


if __name__ == "__main__":
    sys.exit(main())
