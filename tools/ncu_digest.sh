#!/bin/bash
# Turn one `ncu --set full` report into the small text / JSON files that are worth carrying back from the GPU box
# (gpurun copies gpurun_out/ back only while it stays under 64 MiB; a report with source is 25-35 MB):
#   tools/ncu_digest.sh <report.ncu-rep> <level> <tag> <build>
# -> gpurun_out/<tag>_summary.txt (tools/ncu_summary.py), <tag>_lines_<kernel>.txt (stall samples by source line and by
#    function), <tag>_opcodes_<kernel>.txt, gpurun_out/traffic.json (tools/ncu_traffic.py, merged over calls).
# The report itself is kept only when it is smaller than 12 MB.
REP=$1; LVL=$2; TAG=$3; BUILD=${4:-unknown}
LIB=lizard_b200/liblizard_b200.so
python tools/ncu_summary.py $REP > gpurun_out/${TAG}_summary.txt 2>&1
python tools/ncu_traffic.py --level $LVL --build "$BUILD" --out gpurun_out/traffic.json $REP > /dev/null 2>&1
for K in lizard_encode_units lizard_decode_units lizard_decode2_units lizard_huf_expand; do
  if grep -q "$K" gpurun_out/${TAG}_summary.txt; then
    python tools/ncu_lines.py $REP $K $LIB 40 > gpurun_out/${TAG}_lines_$K.txt 2>&1
    python tools/ncu_lines.py $REP $K $LIB 25 func >> gpurun_out/${TAG}_lines_$K.txt 2>&1
    python tools/ncu_opcodes.py $REP $K $LIB > gpurun_out/${TAG}_opcodes_$K.txt 2>&1
  fi
done
SZ=$(stat -c %s $REP 2>/dev/null)
if [ -n "$SZ" ] && [ "$SZ" -gt 12000000 ]; then rm -f $REP; fi
