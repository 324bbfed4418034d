#!/bin/bash
# round-2 single-GPU batch 3 (ping-pong FMHA + converged MMA issuers as defaults): full GPU test suite, N=1 bench line
cd "$(dirname "$0")/../.."
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_e.log 2>&1; tail -6 gpurun_out/r2_gpu_tests_e.log
python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err; tail -c 2500 gpurun_out/r2_bench_e.json | cut -c1-2500; tail -3 gpurun_out/r2_bench_e.err
