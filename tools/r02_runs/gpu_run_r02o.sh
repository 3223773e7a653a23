#!/bin/bash
# r02o: lean (32-bit, record-based) windowed extension chain against the global-memory chain; opaque per-warp shared-memory block pointer
TAG=r02o
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
run base --levels 10,21,41 --variants 7
for v in chain_old win_opq old_opq; do run $v --levels 10,21 --variants 7; done
for v in $EXTRA_VARIANTS; do run $v --levels 10,21 --variants 7; done
el variants
timeout 200 ncu --set full --clock-control none -k regex:lizard_decode_units -s 2 -c 1 -f -o gpurun_out/${TAG}_dec_l10 python tools/ncu_target.py --level 10 --warm 2 2>&1 | tail -1
bash tools/ncu_digest.sh gpurun_out/${TAG}_dec_l10.ncu-rep 10 ${TAG}_dec_l10 "$BUILD"
el ncu
