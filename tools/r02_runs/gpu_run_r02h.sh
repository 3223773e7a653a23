#!/bin/bash
# r02h: L2 policy experiments for the token kernel (evict-first literal loads, evict-last output stores); frame chunk sizes
TAG=r02h
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
for v in nostream streamlits evl; do
  echo "== $v"
  LIZARDB200_LIB=$PWD/lizard_b200/liblizard_b200_$v.so timeout 100 python tools/dec_bench.py --levels 10,21,41 --variants 7 --iters 5 2>&1 | tee -a gpurun_out/${TAG}_policy.log | cut -c1-200
done
el policy
for mib in 64 128 256; do
  LIZARDB200_FRAME_CHUNK_MIB=$mib timeout 150 python bench.py --steps 5 --warmup 3 --legs "" > gpurun_out/${TAG}_bench_chunk$mib.json 2> gpurun_out/${TAG}_bench_chunk$mib.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_chunk$mib.json")); e=d["e2e"]
    print("chunk $mib MiB: e2e", e["value"], "compress_ms", e["compress_ms_rank0"], "decompress_ms", e["decompress_ms_rank0"], "with checksum", e["with_content_checksum"]["value"])
except Exception as ex: print("chunk $mib failed", ex)
PY
done
el chunk-sweep
