#!/bin/bash
# round-2 four-GPU bench line of the final kernels (peer-mapped handoff)
cd "$(dirname "$0")/../.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29841 bench.py --gpus 4 --steps 2 --warmup 3 > gpurun_out/r2_bench_n4_peer.json 2> gpurun_out/r2_bench_n4_peer.err; tail -c 1500 gpurun_out/r2_bench_n4_peer.json; tail -3 gpurun_out/r2_bench_n4_peer.err
