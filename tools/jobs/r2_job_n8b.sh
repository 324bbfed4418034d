#!/bin/bash
# round-2 eight-GPU check of the final kernels: the bench line at N=8 with the default (peer-mapped) handoff
cd "$(dirname "$0")/../.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29821 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/r2_bench_n8_peer.json 2> gpurun_out/r2_bench_n8_peer.err; tail -c 3000 gpurun_out/r2_bench_n8_peer.json; tail -5 gpurun_out/r2_bench_n8_peer.err
