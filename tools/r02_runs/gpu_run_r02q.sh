#!/bin/bash
# r02q: host timeline of the frame round trip (LIZARDB200_TRACE) and the encoder at all levels with the opaque pointers
TAG=r02q
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 200 python tools/e2e_trace.py --trace > gpurun_out/${TAG}_trace.out 2> gpurun_out/${TAG}_trace.err
cat gpurun_out/${TAG}_trace.out; grep "^iter" gpurun_out/${TAG}_trace.err
el trace
timeout 300 python tools/dec_bench.py --iters 5 --levels 10,21,30,41 --variants 7 --encode 2>&1 | grep '^{' | tee gpurun_out/${TAG}_levels.jsonl | cut -c1-200
el levels
timeout 300 python -m pytest tests/test_gpu_encode.py -x -q 2>&1 | tail -3
el pytest-encode
