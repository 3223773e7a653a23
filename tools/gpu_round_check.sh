#!/bin/bash
# One `gpurun` call that checks a build end to end (~2 min on the box):
#   /usr/local/graft/bin/gpurun --timeout 420 -- 'bash tools/gpu_round_check.sh [tag]'
# GPU parity tests, the contract bench lines (level 10 with e2e, levels 21 / 41 kernel-only), the launch list, one
# `ncu --set full` capture holding the encode and the decode kernel (level 10), and a kernel-only sweep of the other levels.
# Everything lands in gpurun_out/<tag>_*; read the captures with tools/ncu_summary.py / ncu_lines.py / ncu_opcodes.py and
# copy what should be judged into profiles/.  Nothing printed by a run under ncu is a bench value.
TAG=${1:-check}
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 240 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.log
el pytest
timeout 100 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_l10.json 2> gpurun_out/${TAG}_bench_l10.err; tail -c 600 gpurun_out/${TAG}_bench_l10.json
el bench10
for lvl in 21 41; do
  timeout 60 python bench.py --steps 3 --warmup 3 --level $lvl --no-e2e > gpurun_out/${TAG}_bench_l$lvl.json 2> gpurun_out/${TAG}_bench_l$lvl.err; tail -c 300 gpurun_out/${TAG}_bench_l$lvl.json
done
el bench21/41
timeout 60 python tools/dec_bench.py --levels 11,30,21,41 --variants 7 --iters 3 --encode 2>&1 | tee gpurun_out/${TAG}_levels.log | cut -c1-160
el levels
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_l10.csv python bench.py --steps 2 --warmup 3 --no-e2e > /dev/null 2>&1
el launches
timeout 150 ncu --set full --import-source on --clock-control none -k regex:lizard_.*_units -s 3 -c 2 -f -o gpurun_out/${TAG}_encdec_l10 python tools/dec_bench.py --levels 10 --variants 7 --iters 1 --encode 2>&1 | tail -1
el ncu
