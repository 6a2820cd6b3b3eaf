timeout -k 10 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q -k "one_launch or long_trajectories or row_mlp" 2>&1 | tail -3
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
