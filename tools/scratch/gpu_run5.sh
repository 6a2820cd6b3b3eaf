mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/test_full5.log
timeout 120 python tools/time_chain.py 2>&1 | tee gpurun_out/chain5.log
MMF_TC_WAIT_HINT=0 timeout 120 python tools/time_chain.py 2>&1 | tee -a gpurun_out/chain5.log
MMF_TC_VARIANT=71 timeout 120 python tools/time_chain.py 2>&1 | tee -a gpurun_out/chain5.log
for sp in 1 3 0.1; do echo "== spread $sp"; SPREAD=$sp timeout 120 python tools/time_resample.py; done 2>&1 | tee gpurun_out/resample5.log
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 3000 gpurun_out/bench_c3.json; tail -5 gpurun_out/bench_c3.err
timeout -k 10 600 python bench.py --workload c1 --steps 5 --warmup 3 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; tail -c 1200 gpurun_out/bench_c1.json; tail -3 gpurun_out/bench_c1.err
