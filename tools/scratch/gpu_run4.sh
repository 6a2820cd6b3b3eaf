mkdir -p gpurun_out
PARITY=1 VARIANTS="72 71 74 41" tools/sweep_variants.sh 2>&1 | tee gpurun_out/sweep2.log
for p in bf16x3; do MMF_LIB=tools/ubench/libmmf_ablate3.so timeout 120 python tools/ws_timeline.py $p; done > gpurun_out/timeline2.log 2>&1
for a in 1 2; do MMF_LIB=tools/ubench/libmmf_ablate$a.so timeout 120 python tools/time_chain.py; done > gpurun_out/ablate2.log 2>&1
cat gpurun_out/ablate2.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "reproducible or heads or bptt" 2>&1 | tail -5
timeout 1500 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -s -k "c2" 2>&1 | tee gpurun_out/baseline_tests2.log | tail -15
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resample_fast -s 2 -c 1 -f -o gpurun_out/prof_resample_fast python tools/time_resample.py > gpurun_out/ncu_rf.log 2>&1; tail -2 gpurun_out/ncu_rf.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_particle_chain_ws -s 2 -c 1 -f -o gpurun_out/prof_chain_ws python tools/time_chain.py bf16x3 > gpurun_out/ncu_ws.log 2>&1; tail -2 gpurun_out/ncu_ws.log
ls -la gpurun_out/*.ncu-rep
