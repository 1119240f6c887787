mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_i3d.py -m gpu -k "pool_backward or cosine" -q --timeout 600 -p no:cacheprovider -rf > gpurun_out/c6_pytest.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/c6_pytest.log
for v in 0 1; do
  echo -n "FAV_POOL_BWD_POOLED=$v: "
  FAV_POOL_BWD_POOLED=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sustained-sec 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"
done
