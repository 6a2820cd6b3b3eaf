mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x -k "reweight_train or bptt or c4_gradients or training_glue or train_e2e or reproducible" 2>&1 | tail -6
timeout -k 10 600 python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 700 gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
