mkdir -p gpurun_out
run() { echo -n "$*: "; env "$@" timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --sustained-sec 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"; }
run FAV_PDL=0
run FAV_PDL=1
run FAV_PDL=0 FAV_BRANCH_STREAMS=0
run FAV_PDL=1 FAV_BRANCH_STREAMS=0
run FAV_PDL=0
