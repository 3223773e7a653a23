timeout 300 python -m pytest tests/test_gpu_decode.py -x -q 2>&1 | tail -6
timeout 200 python tools/dec_bench.py --levels 10,21,41,30 --variants 15,7,3 --iters 5 2>&1 | tee gpurun_out/dec_bench_r2.log | tail -14
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_l41_v15.csv python tools/dec_bench.py --levels 41,10 --variants 15 --iters 1 --size-mib 1024 > /dev/null 2>&1
grep -E "huf|decode|token" gpurun_out/launches_l41_v15.csv | tail -12
timeout 300 ncu --set full --import-source on --clock-control none -k regex:lizard_decode -s 3 -c 1 -f -o gpurun_out/dec_v15_l10 python tools/dec_bench.py --levels 10 --variants 15 --iters 2 2>&1 | tail -2
ls -la gpurun_out/
