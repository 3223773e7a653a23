#!/bin/bash
# r02l: after the literals-prefetch and expand-kernel changes -- parity, all levels, the next-batch prefetch variant, the level-10
# encoder shapes again, fresh ncu digests of the token kernel (level 10) and the expand kernel (level 41)
TAG=r02l
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
BUILD=$(cat .build_id 2>/dev/null)
timeout 400 python -m pytest tests/test_gpu_decode.py -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.log
el pytest
OUT=gpurun_out/${TAG}_variants.jsonl
: > $OUT
run() {  # name, args...
  local name=$1; shift
  local lib=lizard_b200/variants/$name.so
  [ "$name" = base ] && lib=lizard_b200/liblizard_b200.so
  LIZARDB200_LIB=$PWD/$lib timeout 200 python tools/dec_bench.py --iters 5 "$@" 2>&1 | grep '^{' | sed "s/^{/{\"build\": \"$name\", /" | tee -a $OUT | cut -c1-200
}
run base --levels 10,21,30,41,11,20,31,40 --variants 7
run base --levels 10 --no-decode --enc-shapes "14,7,2;14,5,2;14,4,2;14,3,2;14,2,2;14,7,2"
el base
for v in pfnext pf0_64b; do run $v --levels 10,21 --variants 7; done
el token-variants
timeout 200 ncu --set full --clock-control none -k regex:lizard_decode_units -s 2 -c 1 -f -o gpurun_out/${TAG}_dec_l10 python tools/ncu_target.py --level 10 --warm 2 2>&1 | tail -1
bash tools/ncu_digest.sh gpurun_out/${TAG}_dec_l10.ncu-rep 10 ${TAG}_dec_l10 "$BUILD"
timeout 200 ncu --set full --clock-control none -k regex:lizard_huf_expand -s 2 -c 1 -f -o gpurun_out/${TAG}_exp_l41 python tools/ncu_target.py --level 41 --warm 2 2>&1 | tail -1
bash tools/ncu_digest.sh gpurun_out/${TAG}_exp_l41.ncu-rep 41 ${TAG}_exp_l41 "$BUILD"
el ncu
