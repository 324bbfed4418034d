#!/bin/bash
# round-2 single-GPU batch 5 (re-created container): full GPU test suite incl. the fmha_poly checks, the polynomial-share x
# hand-over sweep of the ping-pong FMHA, in-situ A/B of fmha_poly on the full step
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_g.log 2>&1; tail -4 gpurun_out/r2_gpu_tests_g.log
timeout 400 python tools/attn_poly_sweep.py > gpurun_out/r2_attn_poly_sweep.log 2>&1; tail -45 gpurun_out/r2_attn_poly_sweep.log | cut -c1-330
timeout 400 python tools/ab_switch.py fmha_poly=0,3,4,6 --rounds 3 --steps 5 --out ab_switch_fmha_poly.json > gpurun_out/r2_ab_fmha_poly.log 2>&1; tail -6 gpurun_out/r2_ab_fmha_poly.log | cut -c1-400
