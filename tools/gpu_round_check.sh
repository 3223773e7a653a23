#!/bin/bash
# One `gpurun` call that checks a build end to end:
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_round_check.sh [tag] [steps: t b n]'
# t = GPU parity tests, b = the contract bench line (level 10 with e2e + legs 21/41), n = `ncu --set full` captures of the
# encode / expand / decode kernels at levels 10, 21, 41 (tools/ncu_target.py) + the launch list of bench.py.
# Everything lands in gpurun_out/<tag>_*; read the captures with tools/ncu_summary.py / ncu_lines.py / ncu_opcodes.py /
# ncu_traffic.py and copy what should be judged into profiles/.  Nothing printed by a run under ncu is a bench value.
TAG=${1:-check}
WHAT=${2:-tbn}
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader | head -2
if [[ $WHAT == *t* ]]; then
  timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.log
  el pytest
fi
if [[ $WHAT == *b* ]]; then
  timeout 200 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_l10.json 2> gpurun_out/${TAG}_bench_l10.err; tail -c 1500 gpurun_out/${TAG}_bench_l10.json; tail -3 gpurun_out/${TAG}_bench_l10.err
  el bench10
  timeout 120 python bench.py --impl reference --steps 3 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; tail -c 900 gpurun_out/${TAG}_bench_ref.json
  el benchref
fi
if [[ $WHAT == *n* ]]; then
  timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_l10.csv python bench.py --steps 2 --warmup 3 --no-e2e --legs "" > /dev/null 2>&1
  el launches
  for lvl in 10 21 41; do
    timeout 240 ncu --set full --import-source on --clock-control none -k regex:lizard_ -s 8 -c 4 -f -o gpurun_out/${TAG}_l${lvl} python tools/ncu_target.py --level $lvl --warm 2 2>&1 | tail -2
    el ncu$lvl
  done
  ls -la gpurun_out/*.ncu-rep
fi
