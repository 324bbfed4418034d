#!/bin/bash
# round-2 final single-GPU validation (last session): smoke(), full GPU test suite, N=1 bench line, reference arm,
# ncu --set full of the one-launch CLIP attention kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_h.log 2>&1; tail -2 gpurun_out/r2_smoke_h.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_h.log 2>&1; tail -4 gpurun_out/r2_gpu_tests_h.log
python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_h.json 2> gpurun_out/r2_bench_h.err; tail -c 700 gpurun_out/r2_bench_h.json; tail -2 gpurun_out/r2_bench_h.err
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2_bench_ref_h.json 2> gpurun_out/r2_bench_ref_h.err; tail -c 500 gpurun_out/r2_bench_ref_h.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_small -s 2 -c 1 -f -o gpurun_out/r2_attn_small python tools/vae_bench.py --clip-only > gpurun_out/r2_ncu_attn_small.log 2>&1; tail -2 gpurun_out/r2_ncu_attn_small.log | cut -c1-300
