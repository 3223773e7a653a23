#!/bin/bash
# r02f: Huffman expand kernel -- CTAs per SM x first-level bits x prefetch distance; frame chunk 64 MiB; whole-file C host;
# ncu traffic of the decode-side kernels at levels 10 / 21 / 41
TAG=r02f
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
BUILD=$(cat .build_id 2>/dev/null)
for cfg in "l1b9_pf128 1" "l1b9_pf128 2" "l1b9_pf128 3" "l1b9_pf0 2" "l1b9_pf0 3" "l1b8_pf128 2" "l1b8_pf128 3" "l1b10_pf128 1"; do
  set -- $cfg
  echo "== $1 ctas $2"
  LIZARDB200_LIB=$PWD/lizard_b200/liblizard_b200_$1.so LIZARDB200_EXP_CTAS_PER_SM=$2 timeout 100 python tools/dec_bench.py --levels 41 --variants 7 --iters 5 2>&1 | tee -a gpurun_out/${TAG}_expand.log | cut -c1-200
done
el expand-variants
LIZARDB200_FRAME_CHUNK_MIB=64 timeout 150 python bench.py --steps 5 --warmup 3 --legs "" > gpurun_out/${TAG}_bench_chunk64.json 2> gpurun_out/${TAG}_bench_chunk64.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_chunk64.json")); e=d["e2e"]
print("chunk 64 MiB: e2e", e["value"], "compress_ms", e["compress_ms_rank0"], "decompress_ms", e["decompress_ms_rank0"], "with checksum", e["with_content_checksum"]["value"])
PY
el chunk64
timeout 200 python -m pytest tests/test_gpu_c_host.py -x -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest_chost.log
el chost
for lvl in 10 21 41; do
  timeout 240 ncu --set full --clock-control none -k "regex:lizard_decode_units|lizard_huf_expand" -s 4 -c 2 -f -o gpurun_out/${TAG}_dec_l${lvl} python tools/ncu_target.py --level $lvl --warm 2 2>&1 | tail -1
  bash tools/ncu_digest.sh gpurun_out/${TAG}_dec_l${lvl}.ncu-rep $lvl ${TAG}_dec_l${lvl} "$BUILD"
  el ncu-dec$lvl
done
timeout 200 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_l10.json 2> gpurun_out/${TAG}_bench_l10.err; tail -c 400 gpurun_out/${TAG}_bench_l10.json
el bench
du -sh gpurun_out
