#!/bin/bash
# r02j: second-generation decode kernel after the mechanical pass (slot-based records, branch-free search, sleeping waits, cold
# paths out of line, 32-bit offsets)
TAG=r02j
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
BUILD=$(cat .build_id 2>/dev/null)
timeout 300 python -m pytest tests/test_gpu_decode.py -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.log
el pytest
timeout 100 python tools/dec_bench.py --levels 10,21 --variants 7,23 --iters 5 2>&1 | tee gpurun_out/${TAG}_dec.log | cut -c1-200
LIZARDB200_DEC2_STAGES=8 timeout 100 python tools/dec_bench.py --levels 10,21 --variants 23 --iters 5 2>&1 | tee gpurun_out/${TAG}_dec_st8.log | cut -c1-200
el dec_bench
timeout 200 ncu --set full --clock-control none -k regex:lizard_decode2 -s 2 -c 1 -f -o gpurun_out/${TAG}_dec2_l10 python tools/ncu_target.py --level 10 --warm 2 --variant 23 2>&1 | tail -1
bash tools/ncu_digest.sh gpurun_out/${TAG}_dec2_l10.ncu-rep 10 ${TAG}_dec2_l10 "$BUILD"
el ncu
