#!/bin/bash
# One gpurun call: GPU parity tests, decode-variant A/B, contract bench lines, launch list.  Outputs under gpurun_out/.
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r3_pytest.log
el pytest done
timeout 200 python tools/dec_bench.py --levels 10,41,21,30 --variants 15,7,11,3 --iters 5 2>&1 | tee gpurun_out/r3_dec_bench.log | tail -20
el dec_bench done
timeout 240 python bench.py --steps 5 --warmup 3 > gpurun_out/r3_bench_l10.json 2> gpurun_out/r3_bench_l10.err
tail -c 2500 gpurun_out/r3_bench_l10.json
el bench l10 done
timeout 160 python bench.py --steps 3 --warmup 3 --level 41 --no-e2e > gpurun_out/r3_bench_l41.json 2> gpurun_out/r3_bench_l41.err
tail -c 1200 gpurun_out/r3_bench_l41.json
el bench l41 done
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3_launches_l10.csv python bench.py --steps 2 --warmup 3 --no-e2e > /dev/null 2>&1
grep -c lizard gpurun_out/r3_launches_l10.csv
el launch list done
