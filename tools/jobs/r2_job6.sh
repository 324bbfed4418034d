#!/bin/bash
# round-2 single-GPU batch 6: one-launch short-sequence attention (svdpp_attn_small_f16) and the CLIP image encoder on it
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet.py -q -k "attn_small or clip" > gpurun_out/r2_gpu_tests_clip.log 2>&1; tail -25 gpurun_out/r2_gpu_tests_clip.log | cut -c1-600
timeout 600 python tools/vae_bench.py --clip-only > gpurun_out/r2_clip_bench.log 2>&1; tail -3 gpurun_out/r2_clip_bench.log | cut -c1-2500
