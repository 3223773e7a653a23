#!/bin/bash
# r02g: Huffman expand kernel v2 (symbol/length byte tables, 16-byte stores, register window for the bitstream)
TAG=r02g
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
BUILD=$(cat .build_id 2>/dev/null)
timeout 300 python -m pytest tests/test_gpu_decode.py tests/test_gpu_frame.py -x -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.log
el pytest
timeout 120 python tools/dec_bench.py --levels 41,30,40 --variants 7,3 --iters 5 2>&1 | tee gpurun_out/${TAG}_dec.log | cut -c1-200
el dec_bench
LIZARDB200_EXP_CTAS_PER_SM=1 timeout 240 ncu --set full --clock-control none -k "regex:lizard_huf_expand" -s 2 -c 1 -f -o gpurun_out/${TAG}_exp_l41 python tools/ncu_target.py --level 41 --warm 2 2>&1 | tail -1
bash tools/ncu_digest.sh gpurun_out/${TAG}_exp_l41.ncu-rep 41 ${TAG}_exp_l41 "$BUILD"
el ncu
