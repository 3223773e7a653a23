#!/usr/bin/env python
"""Merge the DRAM traffic of the kernels in `ncu --set full` reports into a JSON file that bench.py reads for
`roofline.traffic` (so the figure is reproducible from a committed capture instead of typed into bench.py).

  python tools/ncu_traffic.py --level 10 --build <git sha / tag> --out profiles/r02_traffic.json gpurun_out/x.ncu-rep [...]

Works without a GPU (ncu -i).  For every kernel name the LAST launch in the report is taken (the captures skip the warm-up
launches with -s).  Entry: {"<level>": {"<kernel>": {"traffic": read+write bytes, "dram_read": .., "dram_write": ..,
"duration_ms": .., "report": file, "build": ..}}}.
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, required=True)
    ap.add_argument("--build", default="")
    ap.add_argument("--out", required=True)
    ap.add_argument("reports", nargs="+")
    a = ap.parse_args()
    data = {}
    if os.path.exists(a.out):
        with open(a.out) as f:
            data = json.load(f)
    lvl = data.setdefault(str(a.level), {})
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    tscale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}
    for rep in a.reports:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            raise SystemExit("no kernels in " + rep)
        hdr, units = rows[0], rows[1]
        u = dict(zip(hdr, units))
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            m = re.search(r"(lizard_\w+)", d["Kernel Name"])        # "void <unnamed>::lizard_decode_units_kernel<3>(...)" -> the bare name
            name = m.group(1) if m else d["Kernel Name"].split("(")[0].strip()
            rd = float(d["dram__bytes_read.sum"].replace(",", "")) * scale.get(u["dram__bytes_read.sum"], 1.0)
            wr = float(d["dram__bytes_write.sum"].replace(",", "")) * scale.get(u["dram__bytes_write.sum"], 1.0)
            dur = float(d["gpu__time_duration.sum"].replace(",", "")) * tscale.get(u["gpu__time_duration.sum"], 1.0)
            lvl[name] = {"traffic": rd + wr, "dram_read": rd, "dram_write": wr, "duration_ms": round(dur, 4),
                         "report": os.path.basename(rep), "build": a.build,
                         "grid": d.get("launch__grid_size", ""), "block": d.get("launch__block_size", "")}
    with open(a.out, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)
        f.write("\n")
    print(json.dumps(lvl, indent=1))


if __name__ == "__main__":
    main()
