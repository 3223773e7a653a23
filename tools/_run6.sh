#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
run() { timeout 100 python tools/dec_bench.py --levels $2 --variants 7 --iters 5 2>&1 | tee gpurun_out/r6_$1.log | cut -c1-140 | tail -4; el $1; }
cp lizard_b200/liblizard_b200.so /tmp/keep.so
run default_l1_9 41,30,10,21
LIZARDB200_EXP_CARVEOUT=-1 LIZARDB200_DEC_CARVEOUT=-1 run default_l1_9_drivercarve 41,10
cp lizard_b200/_variant_tabglobal.so lizard_b200/liblizard_b200.so
run tabglobal_l1_9 41,30,10,21
timeout 100 python -m pytest tests/test_gpu_decode.py -x -q 2>&1 | tail -2
cp lizard_b200/_variant_l1_10.so lizard_b200/liblizard_b200.so
run l1_10 41,30
cp lizard_b200/_variant_l1_8.so lizard_b200/liblizard_b200.so
run l1_8_tabglobal 41,30,10
cp /tmp/keep.so lizard_b200/liblizard_b200.so
