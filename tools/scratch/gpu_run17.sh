mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "resample or normalize" 2>&1 | tail -4
{ timeout 300 python tools/profile_c5.py 1048576; timeout 300 python tools/profile_c5.py 262144; timeout 300 python tools/profile_c5.py 65536; timeout 300 python tools/profile_c5.py 1048576 128 multinomial_fast; timeout 300 python tools/profile_c5.py 1048576 128 systematic; } 2>&1 | grep -v Warning | grep "M=\|normalize" | tee gpurun_out/profile_c5_big.log
