"""Writes profiles/<tag>_<kernel>.txt from an ncu report: key raw metrics, opcode mix, stall mix.
Usage: python scripts/profile_summary.py <report.ncu-rep> <out.txt> [title]"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
with open(out, "w") as f:
    f.write(f"# {title}\n# source: ncu --set full --clock-control none --import-source on (one launch, B200)\n\n")
    for h, u, v in zip(hdr, units, vals):
        if h in want:
            f.write(f"{h:70s} {v:>22s} {u}\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    summ = subprocess.run([sys.executable, "scripts/ncu_source_summary.py", "24"], input=src, capture_output=True, text=True).stdout
    f.write("\n## executed warp-instructions by SASS opcode, stall-sample mix\n" + summ)
print(open(out).read())
