mkdir -p gpurun_out
timeout -k 10 2400 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/test_full6.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke6.log
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 1500 gpurun_out/bench_c3.json; tail -5 gpurun_out/bench_c3.err
timeout -k 10 600 python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 800 gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
timeout -k 10 600 python bench.py --workload c1 --steps 5 --warmup 3 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; tail -c 600 gpurun_out/bench_c1.json; tail -3 gpurun_out/bench_c1.err
timeout -k 10 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck_smoke.log 2>&1; tail -6 gpurun_out/sanitizer_racecheck_smoke.log
timeout -k 10 900 compute-sanitizer --tool initcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_initcheck_smoke.log 2>&1; tail -6 gpurun_out/sanitizer_initcheck_smoke.log
