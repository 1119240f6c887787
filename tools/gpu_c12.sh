mkdir -p gpurun_out
run() { echo -n "$*: "; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sustained-sec 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items() if 'pool' in k})"; }
run FAV_POOL_THREADS=800
run FAV_POOL_THREADS=400 FAV_POOL_CGN=4
run FAV_POOL_THREADS=400 FAV_POOL_CGN=2
run FAV_POOL_THREADS=384 FAV_POOL_CGN=2
run FAV_POOL_THREADS=512 FAV_POOL_CGN=2
run FAV_POOL_THREADS=800 FAV_POOL_CGN=2
