#!/bin/bash
# r02b: first GPU contact of the second-generation decode kernel
TAG=r02b
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
BUILD=$(cat .build_id 2>/dev/null)
timeout 400 python -m pytest tests/test_gpu_decode.py -x -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest_decode.log
el pytest-decode
timeout 120 python tools/dec_bench.py --levels 10,21,41 --variants 7,23 --iters 5 2>&1 | tee gpurun_out/${TAG}_dec.log | cut -c1-200
el dec_bench
for st in 8; do
  LIZARDB200_DEC2_STAGES=$st timeout 100 python tools/dec_bench.py --levels 10,21 --variants 23 --iters 5 2>&1 | tee gpurun_out/${TAG}_dec_st$st.log | cut -c1-200
done
for c in 8 10 12; do
  LIZARDB200_DEC2_CTAS_PER_SM=$c timeout 100 python tools/dec_bench.py --levels 10 --variants 23 --iters 5 2>&1 | tee gpurun_out/${TAG}_dec_ctas$c.log | cut -c1-200
done
el sweeps
timeout 200 ncu --set full --clock-control none -k regex:lizard_decode2 -s 2 -c 1 -f -o gpurun_out/${TAG}_dec2_l10 python tools/ncu_target.py --level 10 --warm 2 2>&1 | tail -2
bash tools/ncu_digest.sh gpurun_out/${TAG}_dec2_l10.ncu-rep 10 ${TAG}_dec2_l10 "$BUILD"
el ncu-dec2
for lvl in 10 21 41; do
  timeout 240 ncu --set full --clock-control none -k regex:lizard_encode_units -s 2 -c 1 -f -o gpurun_out/${TAG}_enc_l${lvl} python tools/ncu_target.py --level $lvl --warm 2 2>&1 | tail -1
  bash tools/ncu_digest.sh gpurun_out/${TAG}_enc_l${lvl}.ncu-rep $lvl ${TAG}_enc_l${lvl} "$BUILD"
  el ncu-enc$lvl
done
timeout 300 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_decode.py 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest_rest.log
el pytest-rest
du -sh gpurun_out
