run() { lvl=$1; shape=$2; LIZARDB200_ENC_SHAPE=$shape timeout 150 python bench.py --level $lvl --steps 2 --warmup 3 --no-e2e --cpu-sample-mib 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('level $lvl shape $shape', d['config']['compress_MBps'], d['config']['decompress_MBps'])"; }
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run 10 default
run 10 1,1,24
run 10 12,12,2
run 10 14,13,1
run 21 default
run 21 1,1,6
run 21 14,3,2
run 21 8,3,2
run 41 default
run 41 1,1,5
run 30 default
run 30 1,1,16
