mkdir -p gpurun_out
{ for t in 1 4096 2048; do echo "== MMF_RESAMPLE_BIG=$t"; for M in 16384 4096; do MMF_RESAMPLE_BIG=$t timeout 300 python tools/profile_c5.py $M; done; done; } 2>&1 | grep -v Warning | grep "==\|M=\|normalize" | tee gpurun_out/profile_c5_thresh.log
timeout -k 10 900 python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c5.json').read().strip().splitlines()[-1]); print('C5', d['value'], [(x['particles'], round(x['particle_steps_per_s']/1e9,3)) for x in d['sweep']])"
