#!/bin/bash
# round-2 two-GPU bench line of the final tree (peer-mapped handoff)
cd "$(dirname "$0")/../.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29851 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err; tail -c 900 gpurun_out/r2_bench_n2_final.json; tail -2 gpurun_out/r2_bench_n2_final.err
