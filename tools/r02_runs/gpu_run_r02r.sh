#!/bin/bash
# r02r: host timeline of the frame round trip with the decoder's chunk ramp; chunk size sweep under the ramp
TAG=r02r
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 200 python tools/e2e_trace.py --trace > gpurun_out/${TAG}_trace.out 2> gpurun_out/${TAG}_trace.err
cat gpurun_out/${TAG}_trace.out; grep "^iter" gpurun_out/${TAG}_trace.err
el trace
for mib in 32 64 128; do
  echo "chunk $mib MiB"
  LIZARDB200_FRAME_CHUNK_MIB=$mib timeout 100 python tools/e2e_trace.py 2>&1 | grep "^iter" | tail -3
done | tee gpurun_out/${TAG}_chunks.log
el chunks
