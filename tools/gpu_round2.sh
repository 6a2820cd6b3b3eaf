#!/bin/bash
# 2-GPU runs: C3 forward (weak scaling, no collective) and C4 BPTT step (gradient all-reduce over NCCL)
mkdir -p gpurun_out
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 1800 gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_2gpu.json 2> gpurun_out/bench_c4_2gpu.err; tail -c 1200 gpurun_out/bench_c4_2gpu.json; tail -3 gpurun_out/bench_c4_2gpu.err
