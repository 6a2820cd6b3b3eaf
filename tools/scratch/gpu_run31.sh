timeout 600 python -m pytest tests -m gpu -q -x -k "resample or normalize or long_traj or one_launch" 2>&1 | tail -2
timeout 300 python tools/profile_c5.py 1048576 2>&1 | grep "M=\|normalize"
