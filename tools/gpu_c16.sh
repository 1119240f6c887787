mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_i3d.py tests/test_gpu_resnet.py tests/test_gpu_eval.py tests/test_gpu_api.py -m gpu -q --timeout 600 -p no:cacheprovider -rf -x > gpurun_out/c16_pytest.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/c16_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
for v in 0 1; do echo -n "FAV_APPLY_LUT=$v: "; FAV_APPLY_LUT=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sustained-sec 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), {k: round(v['ms_per_step'],4) for k,v in d['kernels'].items() if k in ('apply','other')}, d['kernels']['apply'])"; done
FAV_APPLY_LUT=1 timeout 300 python bench.py --config c4 --steps 20 --warmup 3 --no-cpu-baseline --sustained-sec 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4', round(d['ms_per_step'],3), d['kernels']['apply'])"
