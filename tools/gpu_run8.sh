mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
filt() { grep -v "^ *File\|^ *\^\|^\[rank.\]:   File" "$1" | tail -${2:-8}; }
timeout -k 10 200 $TR bench.py --gpus 2 --workload c4 --steps 10 --warmup 3 > gpurun_out/bench_c4_2gpu.json 2> gpurun_out/bench_c4_2gpu.err; tail -c 1200 gpurun_out/bench_c4_2gpu.json; filt gpurun_out/bench_c4_2gpu.err
MMF_NCCL_IN_GRAPH=1 MMF_BENCH_WATCHDOG_S=90 timeout -k 10 150 $TR bench.py --gpus 2 --workload c4 --steps 10 --warmup 3 > gpurun_out/bench_c4_2gpu_ingraph.json 2> gpurun_out/bench_c4_2gpu_ingraph.err; tail -c 1200 gpurun_out/bench_c4_2gpu_ingraph.json; grep -B2 -A12 "most recent call first" gpurun_out/bench_c4_2gpu_ingraph.err | head -60
timeout -k 10 400 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_c3_2gpu.json 2> gpurun_out/bench_c3_2gpu.err; tail -c 2500 gpurun_out/bench_c3_2gpu.json; filt gpurun_out/bench_c3_2gpu.err
tools/ubench/ss_bench > gpurun_out/ss_bench.log 2>&1; cat gpurun_out/ss_bench.log
timeout 600 python -m pytest tests -m gpu -q -x -k "kf_fuse_measurements or measurement_level or row_mlp" 2>&1 | tail -5
