#!/bin/bash
# final verification of the committed state: GPU tests, contract bench lines, launch list, one full capture of the decode kernel
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r9_pytest.log
el pytest
timeout 100 python bench.py --steps 5 --warmup 3 > gpurun_out/r9_bench_l10.json 2> gpurun_out/r9_bench_l10.err; tail -c 600 gpurun_out/r9_bench_l10.json
el bench10
timeout 60 python bench.py --steps 3 --warmup 3 --level 41 --no-e2e > gpurun_out/r9_bench_l41.json 2> gpurun_out/r9_bench_l41.err; tail -c 300 gpurun_out/r9_bench_l41.json
el bench41
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r9_launches_l10.csv python bench.py --steps 2 --warmup 3 --no-e2e > /dev/null 2>&1
el launches
timeout 150 ncu --set full --import-source on --clock-control none -k regex:lizard_.*_units -s 3 -c 3 -f -o gpurun_out/r9_encdec_l10 python tools/dec_bench.py --levels 10 --variants 7 --iters 1 --encode 2>&1 | tail -1
el ncu
