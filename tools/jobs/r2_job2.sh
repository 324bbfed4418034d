#!/bin/bash
# round-2 single-GPU batch 2: VAE / CLIP parity checks, full-size front/back-end timing, FMHA stagger sweep in situ
cd "$(dirname "$0")/../.."
python -m pytest tests/test_gpu_unet.py tests/test_gpu_kernels.py -q -k "vae or clip or softmax or transpose or time_conv or handoff or rev_ or zigzag" > gpurun_out/r2_gpu_tests_d.log 2>&1; tail -8 gpurun_out/r2_gpu_tests_d.log
timeout 900 python tools/vae_bench.py > gpurun_out/r2_vae_bench.log 2>&1; tail -3 gpurun_out/r2_vae_bench.log | cut -c1-1500
