#!/bin/bash
# r02e: full GPU suite (fastBig, checksum threads), frame chunk size sweep, encoder launch shapes at levels 21 / 41
TAG=r02e
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest.log
el pytest
for mib in 8 4 16 32; do
  LIZARDB200_FRAME_CHUNK_MIB=$mib timeout 150 python bench.py --steps 5 --warmup 3 --legs "" > gpurun_out/${TAG}_bench_chunk$mib.json 2> gpurun_out/${TAG}_bench_chunk$mib.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_chunk$mib.json"))
    e=d["e2e"]; print("chunk $mib MiB: e2e", e["value"], "compress_ms", e["compress_ms_rank0"], "decompress_ms", e["decompress_ms_rank0"], "with checksum", e["with_content_checksum"]["value"], "value", d["value"])
except Exception as ex: print("chunk $mib failed", ex)
PY
done
el chunk-sweep
timeout 150 python tools/dec_bench.py --levels 21 --no-decode --encode --enc-shapes "14,2,2;12,2,2;10,2,2;8,2,2;8,3,2;10,3,2;6,3,2;14,1,2;14,0,2" --iters 3 2>&1 | tee gpurun_out/${TAG}_enc_shapes_l21.log | cut -c1-160
timeout 150 python tools/dec_bench.py --levels 41 --no-decode --encode --enc-shapes "14,1,2;12,1,2;10,1,2;8,1,2;8,2,2;10,2,2;14,0,2" --iters 3 2>&1 | tee gpurun_out/${TAG}_enc_shapes_l41.log | cut -c1-160
timeout 100 python tools/dec_bench.py --levels 20,40 --variants 7 --encode --iters 3 2>&1 | tee gpurun_out/${TAG}_l20_40.log | cut -c1-200
el enc-shapes
