mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "resample or normalize or c3_full or full_size" 2>&1 | tail -3
{ echo "== prefetch on"; timeout 120 python tools/time_resample.py; echo "== prefetch off"; MMF_RESAMPLE_PREFETCH=0 timeout 120 python tools/time_resample.py; } 2>&1 | tee gpurun_out/resample_prefetch.log
