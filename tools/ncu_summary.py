#!/usr/bin/env python
"""Summarise an .ncu-rep (run here, no GPU needed): headline metrics per kernel + the CUDA source
lines with the most executed instructions / stall samples.  Usage: ncu_summary.py REPORT [TOPN]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 12
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__grid_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
names = [r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "") for r in data]
print("kernels:", names)
for m in METRICS:
    if m in hdr:
        i = hdr.index(m)
        print(f"{m} [{units[i]}] = {[r[i] for r in data]}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
kernel = None
fpath = ""
cols = None
agg = {}
for line in csv.reader(io.StringIO(src)):
    if not line:
        continue
    if line[0] == "File Path":
        fpath = line[1].rsplit("/", 1)[-1]
        continue
    if line[0] == "Function Name":
        kernel = line[1].split("(")[0].replace("void ", "")
        agg.setdefault(kernel, defaultdict(lambda: [0, 0, 0]))  # inst, thread inst, samples
        continue
    if line[0] == "Line No":
        cols = {}
        for k, name in enumerate(line):
            cols.setdefault(name, k)
        continue
    if cols is None or kernel is None or not line[0].strip().isdigit():
        continue   # SASS rows (empty line number) are already folded into their source line's row
    try:
        inst = int(line[cols["Instructions Executed"]])
        tinst = int(line[cols["Thread Instructions Executed"]])
        smp = int(line[cols["# Samples"]])
    except (ValueError, KeyError, IndexError):
        continue
    a = agg[kernel][f"{fpath}:{line[0]}  {line[1].strip()}"]
    a[0] += inst; a[1] += tinst; a[2] += smp
for kernel, lines in agg.items():
    tot_i = sum(v[0] for v in lines.values()) or 1
    tot_s = sum(v[2] for v in lines.values()) or 1
    print(f"--- {kernel}: {tot_i} warp instructions, {tot_s} samples; top lines by instructions")
    for key, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:topn]:
        print(f"  {100*v[0]/tot_i:5.1f}% inst  {100*v[2]/tot_s:5.1f}% smp  thr/inst {v[1]/max(v[0],1):4.1f}  {key[:130]}")
