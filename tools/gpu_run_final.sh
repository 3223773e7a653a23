#!/bin/bash
# One call that validates a build end to end and produces everything profiles/ quotes:
#   GPU tests, smoke, the contract bench line (+ reference arm), launch list, `ncu --set full` digests of every kernel at
#   levels 10 / 21 / 41 (summary, stall samples by line / function, opcode mix, traffic.json).
TAG=${1:-final}
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
BUILD=$(cat .build_id 2>/dev/null)
rm -f gpurun_out/traffic.json
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.log
el pytest
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
el smoke
timeout 250 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
el bench
timeout 120 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; tail -c 500 gpurun_out/${TAG}_bench_ref.json
el benchref
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_l10.csv python bench.py --steps 2 --warmup 3 --no-e2e --legs "" > /dev/null 2>&1
el launches
# full captures (all kernels of one compress + one decompress) cost ~5 GPU-minutes per level: NCU_LEVELS picks the levels
# (default all three); NCU_EXPAND41=1 adds a capture of the expand kernel alone at level 41 (~1 minute)
for lvl in ${NCU_LEVELS-10 21 41}; do
  timeout 400 ncu --set full --clock-control none -k regex:lizard_ -s 8 -c 4 -f -o gpurun_out/${TAG}_l${lvl} python tools/ncu_target.py --level $lvl --warm 2 2>&1 | tail -1
  bash tools/ncu_digest.sh gpurun_out/${TAG}_l${lvl}.ncu-rep $lvl ${TAG}_l${lvl} "$BUILD"
  el ncu$lvl
done
if [ -n "$NCU_EXPAND41" ]; then
  timeout 200 ncu --set full --clock-control none -k regex:lizard_huf_expand -s 2 -c 1 -f -o gpurun_out/${TAG}_exp_l41 python tools/ncu_target.py --level 41 --warm 2 2>&1 | tail -1
  bash tools/ncu_digest.sh gpurun_out/${TAG}_exp_l41.ncu-rep 41 ${TAG}_exp_l41 "$BUILD"
  el ncu-expand41
fi
timeout 200 python tools/dec_bench.py --levels 10,21,41,11,20,30,40 --variants 7 --iters 3 --encode 2>&1 | tee gpurun_out/${TAG}_levels.log | cut -c1-200
el levels
du -sh gpurun_out
