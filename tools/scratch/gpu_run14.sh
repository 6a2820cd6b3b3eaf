mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x -k "reweight_train or bptt or c4_gradients or training_glue or train_e2e or reproducible or train_mode or hoists" 2>&1 | tail -6
timeout -k 10 600 python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c4.json').read().strip().splitlines()[-1]); print('C4', d['value'], d['ms_per_step'], d['config']['eager_ms_per_step'], d['gpu_launches'], d['config']['final_loss'])"; tail -3 gpurun_out/bench_c4.err
