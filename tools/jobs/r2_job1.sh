#!/bin/bash
# round-2 single-GPU batch: full GPU test suite, in-situ A/B of the FMHA variants and of the zigzag traversal, per-shape
# profile, N=1 bench line.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/../.."
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_c.log 2>&1; tail -6 gpurun_out/r2_gpu_tests_c.log
python tools/ab_attn_insitu.py --impls 2,4 --staggers 0,900 > gpurun_out/r2_ab_attn_insitu.log 2>&1; tail -5 gpurun_out/r2_ab_attn_insitu.log
python tools/ab_switch.py zigzag=0,1 --attn-impl-long 4 --out ab_switch_zigzag.json > gpurun_out/r2_ab_switch_zigzag.log 2>&1; tail -3 gpurun_out/r2_ab_switch_zigzag.log
python tools/profile_forward.py > gpurun_out/r2_profile_forward.log 2>&1; tail -2 gpurun_out/r2_profile_forward.log | cut -c1-400
python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; tail -c 1500 gpurun_out/r2_bench_c.json; tail -3 gpurun_out/r2_bench_c.err
