#!/usr/bin/env python3
"""Local/shared/global memory instructions of one kernel by CUDA source line (executed warp-level counts).
usage: tools/ncu_mem_by_line.py source.csv[.gz]|rep.ncu-rep file.o kernel-substring [top]"""
import collections, csv, gzip, io, os, re, subprocess, sys, tempfile
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
src = gzip.open(rep, "rt").read() if rep.endswith(".gz") else (open(rep).read() if rep.endswith(".csv") else
      subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
assert len(lines) == len(data), (len(lines), len(data))
agg = collections.defaultdict(collections.Counter)
tot = collections.Counter()
for ln, r in zip(lines, data):
    ins = re.sub(r"^@!?U?P\d+\s+", "", r[ix['Source']].strip())
    op = ins.split()[0].split(".")[0] if ins else "?"
    if op in ("LDL", "STL", "LDS", "STS", "LDG", "STG"):
        n = int(r[ix['Instructions Executed']])
        w = 2 if ".64" in ins else (4 if ".128" in ins else 1)
        agg[ln][op] += n * w
        tot[op] += n * w
print("32-bit words per warp access, executed:", dict(tot))
key = lambda kv: -(kv[1]["LDL"] + kv[1]["STL"] + kv[1]["LDG"] + kv[1]["STG"])
for ln, v in sorted(agg.items(), key=key)[:top]:
    print("%-28s %s" % ("%s:%d" % ln, " ".join("%s %10d" % (k, v[k]) for k in ("LDL", "STL", "LDG", "STG", "LDS", "STS") if v[k])))
