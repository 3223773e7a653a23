#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 200 python -m pytest tests/test_gpu_decode.py -x -q 2>&1 | tail -4
el pytest decode
timeout 100 python tools/dec_bench.py --levels 41,30,10 --variants 7,3 --iters 5 2>&1 | tee gpurun_out/r5_dec_bench_l1_9.log | tail -8
el l1=9
cp lizard_b200/liblizard_b200.so /tmp/keep.so
for b in 8 10; do
  cp lizard_b200/_variant_l1_$b.so lizard_b200/liblizard_b200.so
  timeout 100 python tools/dec_bench.py --levels 41,30 --variants 7 --iters 5 2>&1 | tee gpurun_out/r5_dec_bench_l1_$b.log | tail -3
  el l1=$b
done
cp /tmp/keep.so lizard_b200/liblizard_b200.so
timeout 150 ncu --set full --import-source on --clock-control none -k regex:lizard_huf_expand -s 2 -c 1 -f -o gpurun_out/r5_expand_l41 python tools/dec_bench.py --levels 41 --variants 7 --iters 1 2>&1 | tail -1
el ncu expand
