#!/bin/bash
# ncu captures of the current build: Huffman expand kernel + token kernel at level 41, token kernel at level 10; launch list at level 41
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 150 ncu --set full --import-source on --clock-control none -k regex:lizard_huf_expand -s 2 -c 1 -f -o gpurun_out/r4_expand_l41 python tools/dec_bench.py --levels 41 --variants 7 --iters 1 2>&1 | tail -2
el expand capture
timeout 150 ncu --set full --import-source on --clock-control none -k regex:lizard_decode -s 2 -c 1 -f -o gpurun_out/r4_dec_l41 python tools/dec_bench.py --levels 41 --variants 7 --iters 1 2>&1 | tail -2
el decode l41 capture
timeout 150 ncu --set full --import-source on --clock-control none -k regex:lizard_decode -s 2 -c 1 -f -o gpurun_out/r4_dec_l10 python tools/dec_bench.py --levels 10 --variants 7 --iters 1 2>&1 | tail -2
el decode l10 capture
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4_launches_l41.csv python bench.py --steps 2 --warmup 3 --no-e2e --level 41 > /dev/null 2>&1
grep -c lizard gpurun_out/r4_launches_l41.csv
el launch list l41
ls -la gpurun_out
