#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) per kernel launch: duration, DRAM bytes, throughputs, occupancy, top stall reasons.
usage: tools/ncu_summary.py file.ncu-rep [--all-stalls]"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_cyc_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "lim_regs"),
    ("launch__occupancy_limit_shared_mem", "lim_smem"),
    ("launch__grid_size", "grid"),
    ("lts__t_sector_hit_rate.pct", "l2_hit"),
    ("smsp__inst_executed.sum", "inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("smsp__inst_executed_op_local_ld.sum", "local_ld"),  # may not exist
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        print("==", name[:90])
        for k, short in KEYS:
            if k in idx:
                print("   %-16s %s %s" % (short, r[idx[k]], units[idx[k]]))
        st = []
        for h, i in idx.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("   stalls: " + ", ".join("%s %.2f" % (n, v) for v, n in st[: (20 if "--all-stalls" in sys.argv else 6)]))
        if "--pipes" in sys.argv:
            for h, i in idx.items():
                if "pipe" in h and "pct" in h:
                    print("   ", h, r[i])


if __name__ == "__main__":
    main()
