#!/usr/bin/env python
"""Aggregate an ncu SASS source page by CUDA source line.

  ncu -i rep.ncu-rep --page source --csv > sass.csv
  cuobjdump -xelf all lib.so ; nvdisasm -g -c x.cubin > dis.txt
  tools/ncu_lines.py sass.csv dis.txt <kernel mangled name> [top]
nvdisasm gives (address -> file:line) for the kernel, ncu gives per-address instruction and stall-sample counts."""
import csv
import re
import sys

sass_csv, dis, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
addr2line = {}
cur, inside = None, False
for ln in open(dis, errors="replace"):
    if ln.startswith(".text."):
        inside = ln.strip().rstrip(":") == ".text." + kern
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(sass_csv)))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
h = rows[hi]
ci, si, ai = h.index("Instructions Executed"), h.index("# Samples"), h.index("Address")
base = None
agg = {}
tot_i = tot_s = 0
for r in rows[hi + 1:]:
    if len(r) <= ci or not r[ci].isdigit():
        continue
    a = int(r[ai], 16) if r[ai].startswith("0x") or re.match(r"^[0-9a-f]+$", r[ai]) else None
    if a is None:
        continue
    if base is None:
        base = a
    key, ins = addr2line.get(a - base, (None, "?"))
    n, s = int(r[ci]), int(r[si]) if r[si].isdigit() else 0
    tot_i += n
    tot_s += s
    e = agg.setdefault(key, [0, 0])
    e[0] += n
    e[1] += s
print("total warp instructions %d, stall samples %d" % (tot_i, tot_s))
for key, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% samples %5.1f%% inst  %s" % (100.0 * s / max(tot_s, 1), 100.0 * n / max(tot_i, 1), key))
