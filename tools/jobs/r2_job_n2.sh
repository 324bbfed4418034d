#!/bin/bash
# round-2 two-GPU batch: bench at N=2 with both handoff transports (bit-equality vs one GPU inside), the reference-style
# benchmark mode with the peer transport, production mode smoke
cd "$(dirname "$0")/../.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29811 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_bench_n2_nccl.json 2> gpurun_out/r2_bench_n2_nccl.err; tail -c 1800 gpurun_out/r2_bench_n2_nccl.json; tail -3 gpurun_out/r2_bench_n2_nccl.err
$TR --master-port 29812 bench.py --gpus 2 --steps 2 --warmup 3 --transport peer > gpurun_out/r2_bench_n2_peer.json 2> gpurun_out/r2_bench_n2_peer.err; tail -c 1800 gpurun_out/r2_bench_n2_peer.json; tail -5 gpurun_out/r2_bench_n2_peer.err
PYTHONPATH=. $TR --master-port 29813 -m src.modes.benchmark --model svd --total-steps 24 --num-samples 6 --warmup-samples 2 --latent-frames 25 --latent-height 72 --latent-width 128 --transport peer --log-level WARNING > gpurun_out/r2_mode_benchmark_peer.log 2>&1; grep BENCHMARK_JSON gpurun_out/r2_mode_benchmark_peer.log | cut -c1-600
PYTHONPATH=. $TR --master-port 29814 -m src.modes.production --total-steps 8 --latent-shape 1 4 14 72 128 --num-samples 3 > gpurun_out/r2_mode_production.log 2>&1; grep -i "final latent norm" gpurun_out/r2_mode_production.log | tail -3
