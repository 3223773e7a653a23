#!/bin/bash
# Build the product library with extra -D flags into lizard_b200/variants/<name>.so (A/B runs: LIZARDB200_LIB=<path> selects
# the build in the Python binding, tools/dec_bench.py and bench.py).  usage: tools/build_variant.sh <name> [-DX=Y ...]
set -e
R="$(cd "$(dirname "$0")/.." && pwd)"
NAME="$1"; shift
mkdir -p "$R/lizard_b200/variants"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Xlinker -Bsymbolic-functions "$@" \
     -o "$R/lizard_b200/variants/$NAME.so" "${SRC:-$R/lizard_b200/csrc/api.cu}"
echo "built $NAME: $*"
