#!/bin/bash
# round-2 ncu evidence of the final kernels (one GPU): launch list of one eager UNet step, DRAM traffic of the GEMM family,
# one --set full capture of the ping-pong FMHA and of a level-0 pair-tile convolution.  Each only after the plain run exits 0.
cd "$(dirname "$0")/../.."
python tools/profile_forward.py --once > gpurun_out/r2_once.log 2>&1 || { tail -5 gpurun_out/r2_once.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv python tools/profile_forward.py --once > gpurun_out/r2_ncu_list.log 2>&1; tail -2 gpurun_out/r2_ncu_list.log
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tc -c 400 --csv --log-file gpurun_out/r2_gemm_dram.csv python tools/profile_forward.py --once > gpurun_out/r2_ncu_dram.log 2>&1; tail -2 gpurun_out/r2_ncu_dram.log
ncu --set full --clock-control none --import-source on -k regex:attn_spatial3 -c 1 -f -o gpurun_out/r2_attn3 python tools/profile_forward.py --once > gpurun_out/r2_ncu_attn3.log 2>&1; tail -2 gpurun_out/r2_ncu_attn3.log
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 20 -c 6 -f -o gpurun_out/r2_gemm6 python tools/profile_forward.py --once > gpurun_out/r2_ncu_gemm6.log 2>&1; tail -2 gpurun_out/r2_ncu_gemm6.log
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches.csv gpurun_out/r2_gemm_dram.csv
