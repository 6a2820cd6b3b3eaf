mkdir -p gpurun_out
NG=${NG:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
filt() { grep -v "^ *File\|^ *\^\|^\[rank.\]:   File" "$1" | tail -${2:-8}; }
timeout -k 10 400 $TR bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/bench_c3_${NG}gpu.json 2> gpurun_out/bench_c3_${NG}gpu.err; tail -c 400 gpurun_out/bench_c3_${NG}gpu.json; filt gpurun_out/bench_c3_${NG}gpu.err 4
timeout -k 10 200 $TR bench.py --gpus $NG --workload c4 --steps 10 --warmup 3 > gpurun_out/bench_c4_${NG}gpu.json 2> gpurun_out/bench_c4_${NG}gpu.err; tail -c 300 gpurun_out/bench_c4_${NG}gpu.json; filt gpurun_out/bench_c4_${NG}gpu.err 4
timeout -k 10 400 $TR bench.py --gpus $NG --workload c5 --steps 3 --warmup 3 > gpurun_out/bench_c5_${NG}gpu.json 2> gpurun_out/bench_c5_${NG}gpu.err; tail -c 300 gpurun_out/bench_c5_${NG}gpu.json; filt gpurun_out/bench_c5_${NG}gpu.err 4
