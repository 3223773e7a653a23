#!/bin/bash
# r02d: encoder bucket prefetch (levels 21/41), Huffman expand kernel with 2 CTAs per SM and smaller first-level tables
TAG=r02d
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 120 python tools/dec_bench.py --levels 10,21,41 --variants 7 --iters 5 --encode 2>&1 | tee gpurun_out/${TAG}_l1b10.log | cut -c1-220
el base
for b in 9 8; do
  for c in 1 2; do
    LIZARDB200_LIB=$PWD/lizard_b200/liblizard_b200_l1b$b.so LIZARDB200_EXP_CTAS_PER_SM=$c timeout 100 python tools/dec_bench.py --levels 41,30 --variants 7 --iters 5 2>&1 | tee gpurun_out/${TAG}_l1b${b}_c$c.log | cut -c1-220
  done
done
el expand-variants
timeout 300 python -m pytest tests/test_gpu_encode.py tests/test_gpu_decode.py -x -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.log
el pytest
