mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rf -s > gpurun_out/c5_pytest.log 2>&1; echo "pytest exit $?"
tail -8 gpurun_out/c5_pytest.log
grep -E "device bytes|vs oracle" gpurun_out/c5_pytest.log
for v in 0 1; do
  echo -n "FAV_POOL_FWD_WIN=$v: "
  FAV_POOL_FWD_WIN=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sustained-sec 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"
done
