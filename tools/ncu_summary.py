#!/usr/bin/env python
""".ncu-rep -> the small JSON summaries kept under profiles/ (runs anywhere ncu is installed; no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [more.ncu-rep ...] > profiles/rNN_ncu_x_summary.json"""
import csv, json, os, subprocess, sys

KEEP = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex.sum", "lts__t_bytes_srcunit_tex.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg", "gpc__cycles_elapsed.max", "smsp__cycles_elapsed.avg.per_second")
out = {}
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units = rows[0], rows[1]
    out[os.path.splitext(os.path.basename(rep))[0]] = [
        {(f"{n} [{units[head.index(n)]}]" if units[head.index(n)] else n): r[head.index(n)] for n in KEEP if n in head}
        for r in rows[2:]]
json.dump(out, sys.stdout, indent=1)
