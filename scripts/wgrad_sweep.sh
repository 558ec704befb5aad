run() { echo "== $1"; env $1 timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"; }
run "X=1"
run "MMTG_WGRAD_GRID=1"
run "MMTG_WGRAD_STREAM=0"
