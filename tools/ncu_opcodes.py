#!/usr/bin/env python3
"""Executed-instruction histogram by opcode class from an .ncu-rep (source page, SASS).  usage: tools/ncu_opcodes.py rep"""
import csv, io, subprocess, sys, collections, re
src = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
cnt = collections.Counter(); tot = 0
for r in rows[2:]:
    ins = r[ix['Source']].strip()
    ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
    op = ins.split()[0].split(".")[0] if ins else "?"
    n = int(r[ix['Instructions Executed']])
    cnt[op] += n; tot += n
print("total warp instructions", tot)
for op, n in cnt.most_common(40):
    print("%-10s %6.2f%% %d" % (op, 100.0 * n / tot, n))
