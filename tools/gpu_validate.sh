mkdir -p gpurun_out
timeout -k 10 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/test_full11.log
MMF_RESAMPLE_BIG=0 timeout -k 10 900 python -m pytest tests -m gpu -q -k "resample or normalize" 2>&1 | tail -2 | tee gpurun_out/test_resample_cta_path.log
MMF_PF_LOOP_SMALL=1 timeout -k 10 900 python -m pytest tests -m gpu -q -k "one_launch" 2>&1 | tail -2 | tee gpurun_out/test_loop_small_forced.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke11.log
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 600 gpurun_out/bench_c3.json; tail -5 gpurun_out/bench_c3.err
for w in c1 c2; do timeout -k 10 600 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -c 300 gpurun_out/bench_$w.json; tail -3 gpurun_out/bench_$w.err; done
timeout -k 10 600 python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 300 gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
timeout -k 10 900 python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 300 gpurun_out/bench_c5.json; tail -3 gpurun_out/bench_c5.err
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 400 gpurun_out/bench_reference.json
MMF_BENCH_ALLOW_SHORT=1 MMF_BENCH_NO_TRAINING=1 timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-200
