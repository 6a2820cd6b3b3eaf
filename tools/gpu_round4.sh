#!/bin/bash
# ncu captures of the remaining kernels: EKF loop + fusion (config C2) and the BPTT kernels (config C4)
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_ekf_loop -s 3 -c 1 -f -o gpurun_out/prof_ekf_loop python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ekf.log 2>&1; tail -1 gpurun_out/ncu_ekf.log | cut -c1-200
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_kf_fuse -s 3 -c 1 -f -o gpurun_out/prof_kf_fuse python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_kf.log 2>&1; tail -1 gpurun_out/ncu_kf.log | cut -c1-200
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_head_chain_bwd -s 20 -c 1 -f -o gpurun_out/prof_bwd python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bwd.log 2>&1; tail -1 gpurun_out/ncu_bwd.log | cut -c1-200
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_heads_dw -s 20 -c 1 -f -o gpurun_out/prof_dw python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_dw.log 2>&1; tail -1 gpurun_out/ncu_dw.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
