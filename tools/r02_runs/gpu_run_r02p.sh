#!/bin/bash
# r02p: first chain window of a batch issued with cp.async at the top of the batch; opaque block pointer by default; encoder
# with opaque scratch / table pointers and without the source prefetch
TAG=r02p
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
run base --levels 10,21,41,30 --variants 7
run old_opq --levels 10,21 --variants 7
run base --levels 10 --encode --no-decode
for v in enc_opq enc_opq2 e_nopf; do run $v --levels 10 --encode --no-decode; done
el variants
timeout 200 ncu --set full --clock-control none -k regex:lizard_decode_units -s 2 -c 1 -f -o gpurun_out/${TAG}_dec_l10 python tools/ncu_target.py --level 10 --warm 2 2>&1 | tail -1
bash tools/ncu_digest.sh gpurun_out/${TAG}_dec_l10.ncu-rep 10 ${TAG}_dec_l10 "$BUILD"
el ncu
