#!/bin/bash
# r02k: (1) Huffman expand kernel with the bitstream window in a per-lane shared-memory ring (uniform refill) against the
# register-window build, and its prefetch variants; (2) literals-stream prefetch distance of the token loops; (3) 7-warp
# decode CTAs; (4) source prefetch distance of the window parser.  Variant builds: tools/build_variant.sh (lizard_b200/variants).
TAG=r02k
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
  LIZARDB200_LIB=$PWD/$lib timeout 150 python tools/dec_bench.py --iters 5 "$@" 2>&1 | grep '^{' | sed "s/^{/{\"build\": \"$name\", /" | tee -a $OUT | cut -c1-200
}
run base --levels 10,21,30,41 --variants 7 --encode
el base
for v in pf0 pf512 pf1k pf2k pf0_24 pfnone dw7; do run $v --levels 10,21 --variants 7; done
el token-variants
for v in old_regwin hpf0 hpf128 hpf1k_l2; do run $v --levels 41,30 --variants 7; done
el expand-variants
for v in e128 e1024; do run $v --levels 10 --encode --no-decode; done
el encoder-variants
timeout 200 ncu --set full --clock-control none -k regex:lizard_huf_expand -s 2 -c 1 -f -o gpurun_out/${TAG}_exp_l41 python tools/ncu_target.py --level 41 --warm 2 2>&1 | tail -1
bash tools/ncu_digest.sh gpurun_out/${TAG}_exp_l41.ncu-rep 41 ${TAG}_exp_l41 "$BUILD"
el ncu
