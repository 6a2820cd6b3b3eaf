mkdir -p gpurun_out
for p in bf16x3 bf16; do MMF_LIB=tools/ubench/libmmf_ablate3.so timeout 120 python tools/ws_timeline.py $p; done > gpurun_out/timeline.log 2>&1
for a in 1 2; do MMF_LIB=tools/ubench/libmmf_ablate$a.so timeout 120 python tools/time_chain.py; done > gpurun_out/ablate.log 2>&1
timeout 120 python tools/time_chain.py >> gpurun_out/ablate.log 2>&1
MMF_TC_VARIANT=41 timeout 120 python tools/time_chain.py >> gpurun_out/ablate.log 2>&1
cat gpurun_out/ablate.log
timeout 600 python -m pytest tests -m gpu -q -x -k "resample or normalize or filter_steps or full_size_step" 2>&1 | tail -5
timeout 120 python tools/time_resample.py 2>&1 | tee gpurun_out/resample_fast.log
MMF_RESAMPLE_FAST_KERNEL=0 timeout 120 python tools/time_resample.py 2>&1 | tee gpurun_out/resample_old.log
timeout 1500 python -m pytest tests/test_gpu_baseline_configs.py tests/test_training_glue.py -m gpu -q -s 2>&1 | tee gpurun_out/baseline_tests.log | tail -40
