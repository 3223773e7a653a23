#!/usr/bin/env python
"""Aggregate an ncu report's per-instruction stall samples by source line.

usage: ncu_lines.py <report.ncu-rep> <kernel-name-substring> <library.so> [top] [inst|func]
Needs the library compiled with -lineinfo.  Works without a GPU (ncu -i, cuobjdump, nvdisasm)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kern, lib = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
by_inst = len(sys.argv) > 5 and sys.argv[5] == "inst"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", "--print-code", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith("\t.section\t.text.") and kern in l][0]
off2line, cur = {}, None
for l in dis[start + 1:]:
    if l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# a report with several launches prints one table per launch: keep the one whose instruction count is this kernel's
# (NCU_SECTION picks among several launches of the same kernel, default the last)
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
sections = [rows[h:(heads[k + 1] if k + 1 < len(heads) else len(rows))] for k, h in enumerate(heads)]
mine = [sec for sec in sections if abs(sum(1 for r in sec[1:] if len(r) > 5) - len(off2line)) <= 2] or sections
sec = mine[int(os.environ.get("NCU_SECTION", "-1"))]
hdr = sec[0]
rows = [None, hdr] + [r for r in sec[1:] if len(r) > 5]
iS, iI, iA = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Address")
base = int(rows[2][iA], 16)
samp, inst = collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) <= iI:
        continue
    k = off2line.get(int(r[iA], 16) - base)
    samp[k] += int(r[iS] or 0)
    inst[k] += int(r[iI] or 0)
ts, ti = sum(samp.values()), sum(inst.values())
if len(sys.argv) > 5 and sys.argv[5] == "func":
    # bucket by enclosing function: the nearest preceding line at column 0 that looks like a definition
    fcache = {}
    def func_of(k):
        if not k or not os.path.exists(k[0]):
            return str(k)
        if k[0] not in fcache:
            starts = []
            for n, line in enumerate(open(k[0]).read().split("\n"), 1):
                if re.match(r"(template\s*<[^>]*>\s*)?(LZ_HD_COLD|LZ_HDM|LZ_HD|LZ_D|__device__|__global__|static|inline|struct)\b.*", line):
                    m = re.search(r"([A-Za-z_0-9:]+)\s*\(", line)
                    starts.append((n, (m.group(1) if m else line.strip()[:40])))
            fcache[k[0]] = starts
        name = "?"
        for n, nm in fcache[k[0]]:
            if n <= k[1]:
                name = nm
            else:
                break
        return os.path.basename(k[0]) + ":" + name
    fs, fi = collections.Counter(), collections.Counter()
    for k in set(samp) | set(inst):
        f = func_of(k)
        fs[f] += samp[k]; fi[f] += inst[k]
    print("kernel %s: %d stall samples, %d warp instructions -- by function" % (kern, ts, ti))
    for f, v in fi.most_common(top):
        print("%5.1f%% samples %5.1f%% inst  %s" % (100.0 * fs[f] / ts, 100.0 * v / ti, f))
    sys.exit(0)
print("kernel %s: %d stall samples, %d warp instructions" % (kern, ts, ti))
cache = {}
for k, v in (inst if by_inst else samp).most_common(top):
    v = samp[k]
    text = ""
    if k and os.path.exists(k[0]):
        cache.setdefault(k[0], open(k[0]).read().split("\n"))
        text = cache[k[0]][k[1] - 1].strip()[:100]
    print("%5.1f%% samples %5.1f%% inst  %s:%s  %s" % (100.0 * v / ts, 100.0 * inst[k] / ti, os.path.basename(k[0]) if k else None, k[1] if k else "", text))
