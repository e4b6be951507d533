#!/usr/bin/env python3
"""Stall samples of one kernel attributed to CUDA source lines: joins the SASS rows of an .ncu-rep (source page) with
the line table of the cubin the kernel came from (nvdisasm -g; needs -lineinfo and the SAME build).
usage: tools/ncu_by_line.py rep.ncu-rep|source.csv[.gz] file.o kernel-substring [top]
(the first function of the cubin whose mangled name contains the substring is taken)"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, obj, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
infn, cur, lines, taken = False, None, [], None
for line in out.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", line)
    if m:
        infn = kern in m.group(1) and taken in (None, m.group(1))
        if infn:
            taken = m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        lines.append(cur)
if rep.endswith(".gz"):
    import gzip
    src = gzip.open(rep, "rt").read()
elif rep.endswith(".csv"):
    src = open(rep).read()
else:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
assert len(lines) == len(data), "cubin does not match the profiled kernel (%d vs %d SASS instructions)" % (len(lines), len(data))
cols = ["stall_long_sb", "stall_barrier", "stall_no_inst", "stall_wait", "stall_short_sb", "stall_math"]
agg = collections.defaultdict(lambda: collections.Counter())
for ln, r in zip(lines, data):
    agg[ln]["n"] += int(r[ix["# Samples"]])
    agg[ln]["exec"] += int(r[ix["Instructions Executed"]])
    for c in cols:
        agg[ln][c] += int(r[ix[c]])
tot = sum(v["n"] for v in agg.values()) or 1
print("samples %d" % tot)
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:top]:
    print("%5.1f%%  %-28s exec %11d  " % (100 * v["n"] / tot, "%s:%d" % ln, v["exec"])
          + " ".join("%s %4.1f%%" % (c[6:], 100 * v[c] / tot) for c in cols))
