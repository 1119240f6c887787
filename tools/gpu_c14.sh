mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rf -x > gpurun_out/c14_pytest.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/c14_pytest.log
for i in 1 2; do timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --sustained-sec 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"; done
timeout 300 python bench.py --config c4 --steps 30 --warmup 3 --no-cpu-baseline --sustained-sec 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4', round(d['ms_per_step'],3), {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"
