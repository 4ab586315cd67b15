#!/usr/bin/env python
"""Summarise `ncu --page raw --csv` files: one line per captured launch with the metrics the
roofline discussion uses.  python tools/ncu_summary.py file_raw.csv [...]"""
import csv
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"),
        ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
        ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "dmma%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("launch__registers_per_thread", "regs"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
        ("launch__occupancy_limit_registers", "occ_reg"), ("launch__occupancy_limit_shared_mem", "occ_smem")]


def main():
    for f in sys.argv[1:]:
        rows = list(csv.reader(open(f)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            if len(r) < len(hdr):
                continue
            parts = [r[idx["Kernel Name"]].split("(")[0].replace("void ", ""), "grid " + r[idx["Grid Size"]].strip(),
                     "block " + r[idx["Block Size"]].strip()]
            for m, short in WANT:
                if m in idx:
                    parts.append(f"{short} {r[idx[m]]} {units[idx[m]]}".strip())
            print(f + ": " + " | ".join(parts))


if __name__ == "__main__":
    main()
