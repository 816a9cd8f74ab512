#!/usr/bin/env python
"""DRAM traffic of the dominant kernels from this round's `ncu --set full` captures -> profiles/r2_ncu_traffic.json,
which bench.py reads for `roofline.traffic` (so the bench line never carries a hand-typed constant).
usage: python profiles/update_traffic.py fp32=gpurun_out/r2a/prof_fp32.ncu-rep bf16=gpurun_out/r2a/prof_bf16.ncu-rep"""
import csv
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1.0)


def main(args):
    table = {}
    path = os.path.join(HERE, "r2_ncu_traffic.json")
    if os.path.exists(path):
        table = json.load(open(path))
    for a in args:
        prec, rep = a.split("=", 1)
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        col = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                         "gpu__time_duration.sum") if k in hdr}
        for r in rows[2:]:
            name = r[col["Kernel Name"]]
            key = None
            m = re.search(r"tc_gemm_persistent_kernel<(?:\(int\))?([01])", name)      # <SPLIT_A[, SINGLE]>
            if m and m.group(1) == "1":
                key = "fwd_" + prec
            elif m and m.group(1) == "0":
                key = "bwd_" + prec
            elif "plm_softmax_kernel" in name:
                key = "softmax_" + prec
            if key is None:
                continue
            table[key] = {
                "kernel": name,
                "dram_bytes_read": to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]),
                "dram_bytes_write": to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]]),
                "duration_under_ncu": r[col["gpu__time_duration.sum"]] + " " + units[col["gpu__time_duration.sum"]],
                "capture": os.path.basename(rep) + ", ncu --set full --clock-control none, config 2",
            }
    json.dump(table, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(table, indent=1, sort_keys=True))


if __name__ == "__main__":
    main(sys.argv[1:])
