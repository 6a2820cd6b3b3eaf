timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "long_trajectories" 2>&1 | grep -v Warn | tail -8
