"""Summarise one kernel of an .ncu-rep as the counters bench.py reports beside its roofline.

    python tools/ncu_counters.py report.ncu-rep env_steps_of_that_launch [kernel-substring] > profiles/x.json
"""
import csv
import io
import json
import subprocess
import sys

rep, steps = sys.argv[1], float(sys.argv[2])
sub = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
row = [r for r in rows[2:] if sub in r[hdr.index("Kernel Name")]][0]
g = lambda k: float(row[hdr.index(k)].replace(",", ""))
inst = g("smsp__inst_executed.sum")
lanes = g("smsp__thread_inst_executed_per_inst_executed.ratio")
out = {
    "kernel": row[hdr.index("Kernel Name")],
    "source": rep.split("/")[-1] + " (ncu --set full --clock-control none)",
    "env_steps_of_the_launch": steps,
    "gpu_time_ms": g("gpu__time_duration.sum") if "ms" in rows[1][hdr.index("gpu__time_duration.sum")] else g("gpu__time_duration.sum") / 1e3,
    "measured_inst_per_step": inst * 32.0 / steps,        # issue slots x 32 lanes per env-step
    "measured_thread_inst_per_step": inst * lanes / steps,  # instructions of active threads per env-step
    "active_lanes": lanes,
    "issue_active": g("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
    "alu_pipe": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active") / 100.0,
    "fma_pipe": g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active") / 100.0,
    "lsu_wavefronts": g("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") / 100.0,
    "warps_active": g("sm__warps_active.avg.pct_of_peak_sustained_active") / 100.0,
    "dram_bytes_read": g("dram__bytes_read.sum"), "dram_bytes_read_unit": rows[1][hdr.index("dram__bytes_read.sum")],
    "dram_bytes_write": g("dram__bytes_write.sum"), "dram_bytes_write_unit": rows[1][hdr.index("dram__bytes_write.sum")],
}
print(json.dumps(out, indent=1))
