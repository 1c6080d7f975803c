"""ncu_summary.py -- the handful of counters DESIGN.md quotes, one row per kernel launch, from an `ncu --set full` report:

    ncu -i report.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv > profiles/rNN_ncu_summary_*.csv
"""
import csv
import sys

COLS = [
    ("Kernel Name", "kernel"),
    ("gpu__time_duration.sum", "time_ms"),
    ("sm__cycles_elapsed.max", "sm_cycles"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct_of_active"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct_of_elapsed"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_operand_fetch_pct"),
    ("sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_elapsed", "shared_pipe_pct"),
    ("l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "smem_bank_reads_pct"),
    ("l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed", "smem_bank_writes_pct"),
    ("sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_elapsed", "tma_pipe_pct"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "fma_pipe_pct"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("smsp__inst_executed.sum", "warp_instructions"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem_per_block"),
]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
out = csv.writer(sys.stdout)
out.writerow([name + (f" [{units[idx[k]]}]" if k in idx and units[idx[k]] else "") for k, name in COLS])
for r in rows[2:]:
    line = []
    for k, _ in COLS:
        v = r[idx[k]] if k in idx else ""
        if k == "Kernel Name":
            v = v.replace("w2x::tc::", "").split("(")[0][:70]
        line.append(v)
    out.writerow(line)
