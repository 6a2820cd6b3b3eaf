#!/bin/bash
# One GPU visit: tests, bench (both precisions), launch list, full ncu capture of the two hot kernels.
set -x
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=60 > gpurun_out/test_full.log 2>&1; tail -4 gpurun_out/test_full.log
timeout -k 10 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err; tail -c 2500 gpurun_out/bench_bf16x3.json
timeout -k 10 600 python bench.py --steps 3 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; tail -c 1200 gpurun_out/bench_bf16.json
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py bf16x3 6 > gpurun_out/ncu_launch.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:k_particle_chain_tc -s 2 -c 1 -f -o gpurun_out/prof_chain_tc python tools/profile_step.py bf16x3 4 > gpurun_out/ncu_tc.log 2>&1; tail -2 gpurun_out/ncu_tc.log
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_normalize_resample -s 2 -c 1 -f -o gpurun_out/prof_resample python tools/profile_step.py bf16x3 4 > gpurun_out/ncu_nr.log 2>&1; tail -2 gpurun_out/ncu_nr.log
ls -la gpurun_out
