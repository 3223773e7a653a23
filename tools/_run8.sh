#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
cp lizard_b200/liblizard_b200.so /tmp/keep.so
timeout 60 python tools/dec_bench.py --levels 41,30 --variants 7 --iters 5 2>&1 | tee gpurun_out/r8_lean.log | cut -c1-140 | tail -2; el lean
cp lizard_b200/_variant_norw.so lizard_b200/liblizard_b200.so
timeout 60 python tools/dec_bench.py --levels 41,30 --variants 7 --iters 5 2>&1 | tee gpurun_out/r8_norw.log | cut -c1-140 | tail -2; el norw
cp /tmp/keep.so lizard_b200/liblizard_b200.so
timeout 60 python tools/dec_bench.py --levels 10 --no-decode --iters 4 --enc-shapes "14,13,2;14,11,2;14,9,2;14,7,2;14,4,2;12,12,2;13,13,2" 2>&1 | tee gpurun_out/r8_enc_l10.log | cut -c1-150; el enc10
timeout 60 python tools/dec_bench.py --levels 21 --no-decode --iters 3 --enc-shapes "14,3,2;14,2,2;14,1,2;8,3,2;10,3,2" 2>&1 | tee gpurun_out/r8_enc_l21.log | cut -c1-150; el enc21
timeout 60 python tools/dec_bench.py --levels 41 --no-decode --iters 3 --enc-shapes "14,1,2;14,2,2;14,0,2;10,2,2" 2>&1 | tee gpurun_out/r8_enc_l41.log | cut -c1-150; el enc41
