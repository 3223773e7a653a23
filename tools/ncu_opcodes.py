#!/usr/bin/env python
"""Executed-instruction mix of one kernel from an ncu report: opcode histogram overall and for a source line range.

usage: ncu_opcodes.py <report.ncu-rep> <kernel-substring> <library.so> [file:lo-hi]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, kern, lib = sys.argv[1:4]
rng = sys.argv[4] if len(sys.argv) > 4 else None
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", "--print-code", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith("\t.section\t.text.") and kern in l][0]
line_of, ins_of, cur = {}, {}, None
for l in dis[start + 1:]:
    if l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
        ins_of[int(m.group(1), 16)] = m.group(2)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# a report with several launches prints one table per launch: keep the one whose instruction count is this kernel's
# (NCU_SECTION picks among several launches of the same kernel, default the last) -- as tools/ncu_lines.py does
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
sections = [rows[h:(heads[k + 1] if k + 1 < len(heads) else len(rows))] for k, h in enumerate(heads)]
mine = [sec for sec in sections if abs(sum(1 for r in sec[1:] if len(r) > 5) - len(ins_of)) <= 2] or sections
sec = mine[int(os.environ.get("NCU_SECTION", "-1"))]
hdr = sec[0]
rows = [None, hdr] + [r for r in sec[1:] if len(r) > 5]
iI, iA = hdr.index("Instructions Executed"), hdr.index("Address")
base = int(rows[2][iA], 16)
f = lo = hi = None
if rng:
    f, r = rng.split(":")
    lo, hi = [int(x) for x in r.split("-")]
opc = collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iI:
        continue
    off = int(r[iA], 16) - base
    n = int(r[iI] or 0)
    ln = line_of.get(off)
    if rng and not (ln and ln[0] == f and lo <= ln[1] <= hi):
        continue
    t = ins_of[off].split()
    op = t[1] if t[0].startswith("@") else t[0]
    opc[op.split(".")[0]] += n
    tot += n
print("total warp instructions%s: %d" % (" in " + rng if rng else "", tot))
for k, v in opc.most_common(30):
    print("%-12s %6.2f%%" % (k, 100.0 * v / tot))
