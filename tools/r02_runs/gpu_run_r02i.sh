#!/bin/bash
# r02i: expand kernel with nibble-length tables; encode traffic at levels 21 / 41 after removing the bucket prefetch
TAG=r02i
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
BUILD=$(cat .build_id 2>/dev/null)
cp profiles/r02_traffic.json gpurun_out/traffic.json
timeout 300 python -m pytest tests/test_gpu_decode.py tests/test_gpu_frame.py tests/test_gpu_encode.py -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.log
el pytest
timeout 120 python tools/dec_bench.py --levels 41,30,40,21 --variants 7 --iters 5 --encode 2>&1 | tee gpurun_out/${TAG}_dec.log | cut -c1-200
el dec_bench
for lvl in 21 41; do
  timeout 200 ncu --set full --clock-control none -k regex:lizard_encode_units -s 2 -c 1 -f -o gpurun_out/${TAG}_enc_l${lvl} python tools/ncu_target.py --level $lvl --warm 2 2>&1 | tail -1
  python tools/ncu_summary.py gpurun_out/${TAG}_enc_l${lvl}.ncu-rep > gpurun_out/${TAG}_enc_l${lvl}_summary.txt 2>&1
  python tools/ncu_traffic.py --level $lvl --build "$BUILD" --out gpurun_out/traffic.json gpurun_out/${TAG}_enc_l${lvl}.ncu-rep > /dev/null 2>&1
  rm -f gpurun_out/${TAG}_enc_l${lvl}.ncu-rep
  el ncu-enc$lvl
done
timeout 150 ncu --set full --clock-control none -k regex:lizard_huf_expand -s 2 -c 1 -f -o gpurun_out/${TAG}_exp_l41 python tools/ncu_target.py --level 41 --warm 2 2>&1 | tail -1
bash tools/ncu_digest.sh gpurun_out/${TAG}_exp_l41.ncu-rep 41 ${TAG}_exp_l41 "$BUILD"
el ncu-exp
