mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q -k "one_launch" 2>&1 | tail -6
MMF_PF_LOOP_SMALL=1 timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "one_launch" 2>&1 | tail -3
{ for pr in bf16x3 fp32; do PREC=$pr timeout 120 python tools/time_loop_small.py; done
PREC=bf16x3 timeout 120 python tools/time_loop_small.py 32 30 50 systematic
PREC=bf16x3 timeout 120 python tools/time_loop_small.py 148 30 50; } 2>&1 | grep -v Warning | tee gpurun_out/loop_small_times2.log
