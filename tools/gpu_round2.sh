#!/bin/bash
set -x
timeout -k 10 600 python -m pytest tests -m gpu -q -k "bptt" > gpurun_out/t_bptt.log 2>&1; tail -5 gpurun_out/t_bptt.log | cut -c1-300
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 1500 gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
