#!/usr/bin/env python
"""DRAM traffic per launch of each kernel in one or more .ncu-rep files (ncu --set full) -> JSON that
bench.py reads for `roofline.traffic`.
usage: python tools/ncu_traffic.py out.json workload=c3 a.ncu-rep [b.ncu-rep ...]"""
import csv
import json
import re
import subprocess
import sys


def main():
    out_path, reps = sys.argv[1], [a for a in sys.argv[2:] if a.endswith(".ncu-rep")]
    meta = dict(a.split("=", 1) for a in sys.argv[2:] if "=" in a and not a.endswith(".ncu-rep"))
    res = {"source": "ncu --set full --clock-control none (dram__bytes_read.sum + dram__bytes_write.sum, per launch)",
           **meta, "kernels": {}}
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}

        def gb(r, name):
            v, u = float(r[idx[name]]), units[idx[name]].lower()
            return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}[u]
        for r in rows[2:]:
            name = re.sub(r"<.*", "", r[idx["Kernel Name"]].replace("void ", "").replace("genpk::", ""))
            res["kernels"][name] = {"dram_read_bytes": gb(r, "dram__bytes_read.sum"), "dram_write_bytes": gb(r, "dram__bytes_write.sum"),
                                    "duration_ms_under_ncu": float(r[idx["gpu__time_duration.sum"]]) *
                                    {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[idx["gpu__time_duration.sum"]].lower().replace("msecond", "ms").replace("usecond", "us").replace("nsecond", "ns").replace("second", "s")],
                                    "report": rep.split("/")[-1]}
    json.dump(res, open(out_path, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
