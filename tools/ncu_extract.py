"""Extracts the judged numbers from the `ncu --set full` captures in gpurun_out/ (scratch) into the tracked profiles/:
  profiles/<round>_ncu_metrics.json : per capture, the handful of counters DESIGN.md and bench.py quote
  profiles/<round>_raw_<capture>.csv : the capture's full raw metric page (one kernel launch)
Usage: python tools/ncu_extract.py r02 prof_chain_ws prof_resample_fast ..."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    "duration": "gpu__time_duration.sum",
    "dram_read": "dram__bytes_read.sum",
    "dram_write": "dram__bytes_write.sum",
    "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "registers": "launch__registers_per_thread",
    "warp_instructions": "smsp__inst_executed.sum",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "tensor_pipe_active_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "lsu_smem_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "l2_throughput_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm_cycles": "sm__cycles_elapsed.max",
}


def main():
    rnd, captures = sys.argv[1], sys.argv[2:]
    path = os.path.join(ROOT, "profiles", f"{rnd}_ncu_metrics.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for cap in captures:
        rep = os.path.join(ROOT, "gpurun_out", cap + ".ncu-rep")
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        header, units, vals = rows[0], rows[1], rows[2]
        with open(os.path.join(ROOT, "profiles", f"{rnd}_raw_{cap}.csv"), "w") as f:
            w = csv.writer(f)
            w.writerow(["metric", "unit", "value"])
            for h, u, v in zip(header, units, vals):
                w.writerow([h, u, v])
        col = {h: i for i, h in enumerate(header)}
        rec = {"kernel": vals[col["Kernel Name"]]}
        for key, metric in WANT.items():
            if metric in col and vals[col[metric]] not in ("", "n/a"):
                try:
                    rec[key] = {"value": float(vals[col[metric]].replace(",", "")), "unit": units[col[metric]]}
                except ValueError:
                    pass
        out[cap] = rec
        print(cap, rec["kernel"], rec.get("duration"))
    with open(path, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
