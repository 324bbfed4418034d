#!/bin/bash
# round-2 final single-GPU validation: smoke(), full GPU test suite, N=1 bench line, reference arm
cd "$(dirname "$0")/../.."
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_f.log 2>&1; tail -2 gpurun_out/r2_smoke_f.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_f.log 2>&1; tail -4 gpurun_out/r2_gpu_tests_f.log
python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err; tail -c 600 gpurun_out/r2_bench_f.json; tail -2 gpurun_out/r2_bench_f.err
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2_bench_ref_f.json 2> gpurun_out/r2_bench_ref_f.err; tail -c 700 gpurun_out/r2_bench_ref_f.json
