#!/usr/bin/env python3
"""SASS instruction count per source file:line bucket for one kernel of a cubin (nvdisasm -g needs -lineinfo).
usage: tools/sass_by_line.py file.cubin kernel-substring [bucket-lines]"""
import collections, re, subprocess, sys
cubin, kern = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
cur, infn, cnt = None, False, collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", line)
    if m:
        infn = kern in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)) // bucket * bucket)
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        cnt[cur] += 1
tot = sum(cnt.values())
print("total", tot, "instr", tot * 16 / 1024, "KB")
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:40]:
    print("%6d %5.1f%%  %s" % (v, 100 * v / tot, k))
