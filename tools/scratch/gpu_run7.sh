mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout -k 10 400 $TR bench.py --gpus 2 --workload c4 --steps 10 --warmup 3 > gpurun_out/bench_c4_2gpu.json 2> gpurun_out/bench_c4_2gpu.err; tail -c 1500 gpurun_out/bench_c4_2gpu.json; grep -v "^ *File\|^ *\^\|^\[rank.\]:   File" gpurun_out/bench_c4_2gpu.err | tail -8
timeout -k 10 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_c3_2gpu.json 2> gpurun_out/bench_c3_2gpu.err; tail -c 2500 gpurun_out/bench_c3_2gpu.json; grep -v "^ *File\|^ *\^\|^\[rank.\]:   File" gpurun_out/bench_c3_2gpu.err | tail -8
