import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
sys.argv = ["bench.py", "--workload", "c4", "--steps", "1", "--warmup", "3"]
os.environ["MMF_BENCH_ALLOW_SHORT"] = "1"
import bench
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    bench.main()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
