mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -s -k "one_launch" 2>&1 | grep -v Warning | tail -12 | tee gpurun_out/test_c1_one_launch.log
SAN="compute-sanitizer --print-limit 5"
K="one_launch_forward_loop and PushCrossmodalParticleFilter-30-multinomial or row_mlp and 37 or kf_fuse_measurements and 257 or heads_backward_kernel or encoder_conv_layers"
timeout -k 10 1200 $SAN --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" > gpurun_out/sanitizer_racecheck_kernels.log 2>&1; tail -6 gpurun_out/sanitizer_racecheck_kernels.log
timeout -k 10 900 $SAN --tool initcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" > gpurun_out/sanitizer_initcheck_kernels.log 2>&1; tail -6 gpurun_out/sanitizer_initcheck_kernels.log
timeout -k 10 900 $SAN --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" > gpurun_out/sanitizer_memcheck_kernels.log 2>&1; tail -6 gpurun_out/sanitizer_memcheck_kernels.log
