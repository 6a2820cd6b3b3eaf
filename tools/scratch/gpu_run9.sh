mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "one_launch" 2>&1 | tail -15 | tee gpurun_out/test_loop_small.log
timeout -k 10 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/test_full9.log
timeout -k 10 600 python bench.py --workload c1 --steps 5 --warmup 3 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; tail -c 1500 gpurun_out/bench_c1.json | head -c 700; tail -3 gpurun_out/bench_c1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_c1.json").read().strip().splitlines()[-1])
print("C1 value %.4g ms/pass %.4f -> us/step %.2f ; e2e ms %.3f dtype %s"%(d["value"], d["ms_per_step"], d["ms_per_step"]*1e3/50, d["e2e"]["ms_per_step"], d["dtype"]))
PY
