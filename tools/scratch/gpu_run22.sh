timeout 600 python -m pytest tests -m gpu -q -x -k "row_mlp or measurement_level or c2_full or ekf or fixture or streams_host" 2>&1 | tail -3
timeout -k 10 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c2.json').read().strip().splitlines()[-1]); print('C2', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"
timeout -k 10 600 python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c1.json').read().strip().splitlines()[-1]); print('C1', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"
MMF_BENCH_ALLOW_SHORT=1 timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_row_mlp -c 12 --csv --log-file gpurun_out/launches_rowmlp.csv python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; grep -o 'k_row_mlp.*' gpurun_out/launches_rowmlp.csv | awk -F'","' '{print $NF}' | head -12 | tr '\n' ' '
