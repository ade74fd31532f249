"""Turns the raw ncu outputs of a round (gpurun_out/) into the tracked summaries under profiles/.

  python profiles/summarize.py launches gpurun_out/r01_launches_bench.csv profiles/r01_launch_summary.csv "<command that was profiled>"
  python profiles/summarize.py full profiles/r01_ncu_full_summary.csv gpurun_out/r01_gibbs_v6.ncu-rep gpurun_out/r01_loo_v4.ncu-rep ...
"""
import csv
import io
import os
import subprocess
import sys
from collections import defaultdict

KEYS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__cycles_active.avg", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def launches(src, dst, note):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        name = r[k].split("(")[0]
        tot[name] += float(r[v].replace(",", "")) * 1e-6
        cnt[name] += 1
    total = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none: {note} (cold-cache, serialised: compare shares with bench.py's CUDA-event stage_ms_per_step)\n")
        f.write("kernel,launches,total_ms,share\n")
        for name in sorted(tot, key=tot.get, reverse=True)[:12]:
            f.write(f"{name},{cnt[name]},{tot[name]:.3f},{tot[name] / total:.4f}\n")


def full(dst, reps):
    with open(dst, "w") as f:
        for rep in reps:
            out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rows = list(csv.reader(io.StringIO(out)))
            hdr, units, vals = rows[0], rows[1], rows[2]
            f.write(f"## {os.path.basename(rep)}\n")
            f.write(f"Kernel Name,,{vals[hdr.index('Kernel Name')]}\n")
            for i, h in enumerate(hdr):
                if h in KEYS or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                    f.write(f"{h},{units[i]},{vals[i]}\n")
            f.write("\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3:])
