#!/bin/bash
# bench (c3) + ncu capture of the encoder conv kernel
mkdir -p gpurun_out
timeout -k 10 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c3_enc.json 2> gpurun_out/bench_c3_enc.err; tail -c 2600 gpurun_out/bench_c3_enc.json
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_enc_conv3x3 -s 5 -c 1 -f -o gpurun_out/prof_enc_conv python tools/debug_enc.py 37 4096 > gpurun_out/ncu_enc.log 2>&1; tail -2 gpurun_out/ncu_enc.log
