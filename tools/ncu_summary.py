#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): one block of headline metrics per captured launch.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN/x_summary.txt"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}: ncu --set full --clock-control none; values per launch")
    for r in rows[2:]:
        print(f"\n{r[idx['Kernel Name']]}  (launch id {r[idx['ID']]})")
        for w in WANT:
            if w in idx:
                print(f"  {w:72s} {r[idx[w]]} {units[idx[w]]}")


if __name__ == "__main__":
    main()
