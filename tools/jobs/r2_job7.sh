#!/bin/bash
# round-2 single-GPU batch 7: where the VAE temporal decoder's time goes (per-call CUDA events, 14- and 11-frame chunks)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python tools/vae_once.py --events --frames 14 > gpurun_out/r2_vae_events_14.log 2>&1; tail -1 gpurun_out/r2_vae_events_14.log | cut -c1-6000
