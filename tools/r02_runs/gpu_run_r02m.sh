#!/bin/bash
# r02m: token-kernel cache experiments: real loads (plain / L1::evict_last) instead of the prefetch hint, second match-source
# line, far L2 prefetch; level-10 encoder at its new default shape
TAG=r02m
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
OUT=gpurun_out/${TAG}_variants.jsonl
: > $OUT
run() {  # name, args...
  local name=$1; shift
  local lib=lizard_b200/variants/$name.so
  [ "$name" = base ] && lib=lizard_b200/liblizard_b200.so
  LIZARDB200_LIB=$PWD/$lib timeout 200 python tools/dec_bench.py --iters 5 "$@" 2>&1 | grep '^{' | sed "s/^{/{\"build\": \"$name\", /" | tee -a $OUT | cut -c1-200
}
run base --levels 10,21 --variants 7 --encode
el base
for v in pfload pfload_el mpf2 pfl2; do run $v --levels 10,21 --variants 7; done
el token-variants
