#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
run() { timeout 100 python tools/dec_bench.py --levels $2 --variants 7 --iters 5 2>&1 | tee gpurun_out/r7_$1.log | cut -c1-140 | tail -4; el $1; }
cp lizard_b200/liblizard_b200.so /tmp/keep.so
run regwin_l1_10 41,30,10,21
LIZARDB200_DEC_CTAS_PER_SM=3 run ctas3 10,21
LIZARDB200_DEC_CTAS_PER_SM=2 run ctas2 10,21
cp lizard_b200/_variant_norw.so lizard_b200/liblizard_b200.so
run noregwin_l1_10 41,30
cp /tmp/keep.so lizard_b200/liblizard_b200.so
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
el pytest
