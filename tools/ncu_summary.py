"""Summarise an .ncu-rep (ncu --set full) into a small text table for profiles/.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

WANT = [
    ("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem_KB"), ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
cols = [(hdr.index(k), n) for k, n in WANT if k in hdr]
print("# source:", sys.argv[1], "(ncu --set full --clock-control none; units as reported by ncu:",
      ", ".join(f"{n}={units[i]}" for i, n in cols if units[i]), ")")
print(" | ".join(n for _, n in cols))
for r in rows[2:]:
    print(" | ".join(r[i][:70] for i, _ in cols))
