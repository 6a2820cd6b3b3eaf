mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "one_launch" 2>&1 | tail -8 | tee gpurun_out/test_loop_small.log
MMF_PF_LOOP_SMALL=1 timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "one_launch" 2>&1 | tail -8 | tee -a gpurun_out/test_loop_small.log
{
for pr in bf16x3 bf16 fp32; do PREC=$pr timeout 120 python tools/time_loop_small.py; done
MMF_PF_LOOP_SMALL=0 timeout 120 python tools/time_loop_small.py
PREC=bf16x3 timeout 120 python tools/time_loop_small.py 148 30 50
MMF_PF_LOOP_SMALL=0 timeout 120 python tools/time_loop_small.py 148 30 50
PREC=bf16x3 timeout 120 python tools/time_loop_small.py 148 64 50
MMF_PF_LOOP_SMALL=0 timeout 120 python tools/time_loop_small.py 148 64 50
MMF_PF_LOOP_SMALL=1 PREC=bf16x3 timeout 120 python tools/time_loop_small.py 148 100 20
} 2>&1 | grep -v Warning | tee gpurun_out/loop_small_times.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pf_loop_small -c 1 -f -o gpurun_out/prof_loop_small python tools/time_loop_small.py > gpurun_out/ncu_ls.log 2>&1; tail -2 gpurun_out/ncu_ls.log
