#!/usr/bin/env python3
"""FP64 operations actually executed (ncu: smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on) against the algorithmic
flop model of bench.py (SURVEY.md 8d asks for exactly this validation).  Reads the gzipped raw-page csv that tools/ncu_extract.sh
leaves behind, and optionally the bench JSON of the same workload for the model figure.
usage: tools/sass_flops.py name=raw.csv.gz[:bench.json[:cells]] ...  > profiles/r02_sass_flops.json"""
import csv, gzip, io, json, sys


def raw(path):
    rows = list(csv.reader(io.StringIO(gzip.open(path, "rt").read())))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    return {a: (str(float(c.replace(",", "")) * scale[u]) if u in scale else c) for a, u, c in zip(rows[0], rows[1], rows[2])}


def num(x):
    return float(x.replace(",", ""))


out = {}
for arg in sys.argv[1:]:
    name, rest = arg.split("=", 1)
    parts = rest.split(":")
    r = raw(parts[0])
    cyc = num(r["smsp__cycles_elapsed.max"])
    d = {op: num(r["smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op]) * cyc for op in ("dfma", "dmul", "dadd")}
    e = {"dfma": d["dfma"], "dmul": d["dmul"], "dadd": d["dadd"], "sass_flops": 2 * d["dfma"] + d["dmul"] + d["dadd"],
         "kernel_ms_under_ncu": num(r["gpu__time_duration.sum"]),
         "fp64_pipe_active_pct": num(r["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]),
         "dram_bytes": (num(r["dram__bytes_read.sum"]) + num(r["dram__bytes_write.sum"])), "source": parts[0]}
    if len(parts) > 1 and parts[1]:
        j = json.loads(open(parts[1]).read().strip().splitlines()[-1])
        e["model_flops"] = j["roofline"]["flops_per_launch"]
        e["sass_over_model"] = e["sass_flops"] / e["model_flops"]
        e["cells"] = j["config"]["cells_per_gpu_per_step"]
    if len(parts) > 2:
        e["cells"] = int(parts[2])
    if "cells" in e:
        e["sass_flops_per_cell"] = e["sass_flops"] / e["cells"]
    out[name] = e
print(json.dumps(out, indent=1))
